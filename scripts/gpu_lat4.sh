cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_mapping.py -m gpu -x -q 2>&1 | tail -8
python scripts/time_latency.py 2>&1 | tail -3
echo "--- graph off"; LL_GRAPH=0 python scripts/time_latency.py 2>&1 | tail -2
