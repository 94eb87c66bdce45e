set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s2_tests.log
timeout 300 python scripts/prof_map.py > gpurun_out/s2_map_prof.json 2> gpurun_out/s2_map_prof.err
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err
tail -3 gpurun_out/s2_tests.log; cat gpurun_out/s2_map_prof.json; tail -c 600 gpurun_out/s2_bench.json
