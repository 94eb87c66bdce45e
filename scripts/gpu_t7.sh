cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/bench_latest.json 2> gpurun_out/bench_latest.err; tail -2 gpurun_out/bench_latest.err
python -c "
import json; d=json.load(open('gpurun_out/bench_latest.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d.get('parity',{}).get('status')); print({k:(v['ms_per_launch'],v['launches_per_step']) for k,v in list(d['roofline']['kernels'].items())[:14]})"
bash scripts/gpu_sanitize.sh r02 2>&1 | grep -E "SUMMARY"
