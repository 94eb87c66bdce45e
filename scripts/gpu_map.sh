cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mapping.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/prof_map.py > gpurun_out/s5_map_prof.json 2> gpurun_out/s5_map_prof.err; cat gpurun_out/s5_map_prof.json
