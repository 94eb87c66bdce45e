cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_latest.json 2> gpurun_out/bench_latest.err
python -c "
import json; d=json.load(open('gpurun_out/bench_latest.json')); print(d['value'], d['ms_per_step'], d['e2e'], d.get('cpu_baseline'), {k:v['ms_per_launch'] for k,v in list(d['roofline']['kernels'].items())[:10]}); print({k:v for k,v in d['roofline'].items() if k!='kernels'})"
