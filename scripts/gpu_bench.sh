cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
if [ -n "$TESTS" ]; then timeout 1500 python -m pytest tests -m gpu -x -q $TESTS 2>&1 | tail -${TAILN:-8}; fi
timeout 900 python bench.py --steps ${STEPS:-20} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_latest.json 2> gpurun_out/bench_latest.err
tail -3 gpurun_out/bench_latest.err
python -c "
import json; d=json.load(open('gpurun_out/bench_latest.json')); print(d['value'], d['ms_per_step'], d['e2e'], d.get('cpu_baseline'), d.get('parity')); print({k:(v['ms_per_launch'],v['launches_per_step']) for k,v in list(d['roofline']['kernels'].items())}); print({k:v for k,v in d['roofline'].items() if k!='kernels'})"
