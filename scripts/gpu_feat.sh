cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_features.py tests/test_gpu_odometry.py -m gpu -x -q 2>&1 | tail -4
python scripts/prof_kernels.py 2>&1 | tail -1
