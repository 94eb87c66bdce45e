cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "mapping or features" 2>&1 | tail -6
timeout 600 python bench.py --config 1 --steps 20 > gpurun_out/r02_config1.json 2> gpurun_out/cfg1.err; tail -2 gpurun_out/cfg1.err; cut -c1-600 gpurun_out/r02_config1.json
timeout 900 python bench.py --config 3 --steps 5 > gpurun_out/r02_config3.json 2> gpurun_out/cfg3.err; tail -3 gpurun_out/cfg3.err; cut -c1-1800 gpurun_out/r02_config3.json
timeout 900 python bench.py --config 4 --stream-scans 1280 --lanes 64 > gpurun_out/r02_config4_n1_s1280.json 2> gpurun_out/cfg4.err; tail -3 gpurun_out/cfg4.err; cut -c1-2500 gpurun_out/r02_config4_n1_s1280.json
