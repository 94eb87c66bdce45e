set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_odometry.py tests/test_gpu_mapping.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/s3_tests.log
tail -5 gpurun_out/s3_tests.log
for d in ${DMAXES:-2 4 8}; do
LL_ASSOC_DMAX=$d timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/s3_bench_d$d.json 2> gpurun_out/s3_bench.err
LL_ASSOC_DMAX=$d LL_DEV_SKIP=16 LL_DEBUG_ASSOC=1 LL_B=8 LL_STEPS=10 python scripts/prof_run.py 2>&1 | tail -2
done
python - <<'P'
import json,os
for n in os.environ.get("DMAXES","2 4 8").split():
    d=json.load(open("gpurun_out/s3_bench_d%s.json"%n))
    print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], {k:v["ms_per_launch"] for k,v in list(d["roofline"]["kernels"].items())[:6]})
P
