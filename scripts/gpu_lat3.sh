cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_odometry.py tests/test_gpu_mapping.py -m gpu -x -q 2>&1 | tail -8
python scripts/time_latency.py 2>&1 | tail -3
LL_LIB_PATH=$PWD/light-loam_b200/build/timing/liblightloam_b200.so python scripts/lm_timing.py 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/b3.json 2>gpurun_out/b3.err; python -c "
import json; d=json.load(open('gpurun_out/b3.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity'))"
