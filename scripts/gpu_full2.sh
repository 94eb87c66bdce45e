cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/time_latency.py 2>&1 | tail -2
timeout 600 python bench.py --config 3 --steps 20 --warmup 3 > gpurun_out/cfg3_sort3.json 2> gpurun_out/cfg3_sort3.err
tail -3 gpurun_out/cfg3_sort3.err; python -c "
import json; d=json.load(open('gpurun_out/cfg3_sort3.json')); print(d['value'], d['ms_per_step'], d['vote_off']['ms_per_step'], d['e2e']); print({k:(v['ms_per_launch'],v['launches_per_step']) for k,v in list(d['roofline']['kernels'].items())[:12]})"
