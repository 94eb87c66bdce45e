cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mapping.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --config 3 --steps 20 --warmup 3 > gpurun_out/cfg3_sort.json 2> gpurun_out/cfg3_sort.err
tail -3 gpurun_out/cfg3_sort.err; cat gpurun_out/cfg3_sort.json | cut -c1-3000
