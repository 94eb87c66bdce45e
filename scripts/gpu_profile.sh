# usage: bash scripts/gpu_profile.sh <tag>   — bench line, ncu launch list of the bench command, ncu --set full of the top kernels
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${1:-r01d}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${TAG}_tests.log; tail -1 gpurun_out/${TAG}_tests.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 400 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
# launch list of the bench command itself (values printed under ncu are not bench values)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 230 -c 92 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_launches.log 2>&1
for k in k_ring_lessflat k_odom_assoc k_ring_sort k_lm_solve_odom k_classify k_odom_assoc_heavy k_scatter k_ring_pick; do
LL_B=${LL_B:-256} LL_STEPS=10 timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$k\$ -s 8 -c 1 -o gpurun_out/${TAG}_$k -f python scripts/prof_run.py > gpurun_out/${TAG}_$k.log 2>&1
done
ls gpurun_out/${TAG}_*.ncu-rep | wc -l
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['value'], d['ms_per_step'], d['e2e'], d.get('cpu_baseline'), {k:v for k,v in d['roofline'].items() if k!='kernels'})"
cat gpurun_out/${TAG}_bench_reference.json | head -c 600
