# final check of a code state: the driver's round-end sequence (GPU tests, smoke, bench both arms) + latency figures
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/final_bench_reference.json 2> gpurun_out/final_bench.err
timeout 900 python bench.py > gpurun_out/final_bench.json 2>> gpurun_out/final_bench.err; tail -c 300 gpurun_out/final_bench.err
python -c "
import json; d=json.load(open('gpurun_out/final_bench.json')); r=json.load(open('gpurun_out/final_bench_reference.json'))
print(d['value'], d['ms_per_step'], d['steps'], d['e2e']['value'], d['parity']['status'], d['gpu_launches'], d['clocks'], d['roofline']['frac'], d['cpu_baseline']['value'], 'ref', r['value'], r['cpu_baseline']['cores'])"
python scripts/time_latency.py 2>&1 | tail -4 | tee gpurun_out/latency.txt
