# usage: bash scripts/gpu_profile.sh <tag>   — bench line, ncu launch list of the bench command, ncu --set full of the top kernels
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${1:-r02}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${TAG}_tests.log; tail -1 gpurun_out/${TAG}_tests.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 400 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
# launch list of the bench command itself (values printed under ncu are not bench values)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-320} -c 64 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/${TAG}_launches.log 2>&1
for k in k_ring_lessflat k_odom_assoc k_ring_sort k_lm_solve_odom k_classify k_odom_assoc_heavy k_scatter k_ring_pick k_odom_vote k_odom_prep k_index_count k_index_scatter k_compact; do
LL_B=${LL_B:-256} LL_STEPS=10 timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$k\$ -s 8 -c 1 -o gpurun_out/${TAG}_$k -f python scripts/prof_run.py > gpurun_out/${TAG}_$k.log 2>&1
done
ls gpurun_out/${TAG}_*.ncu-rep | wc -l
# summaries on the box (gpurun copies back at most 64 MiB): metrics of every report, stall samples / instruction counts per source line
python scripts/ncu_summary.py gpurun_out/${TAG}_launches.csv gpurun_out/${TAG}_k_*.ncu-rep > gpurun_out/${TAG}_ncu_summary.json 2> gpurun_out/${TAG}_ncu_summary.err
for k in k_ring_lessflat k_odom_assoc k_ring_sort k_lm_solve_odom k_classify k_odom_vote k_odom_prep k_index_count; do
python scripts/ncu_lines.py gpurun_out/${TAG}_$k.ncu-rep 25 > gpurun_out/${TAG}_${k}_lines.txt 2>/dev/null
python scripts/ncu_instr.py gpurun_out/${TAG}_$k.ncu-rep 25 > gpurun_out/${TAG}_${k}_instr.txt 2>/dev/null
done
for f in gpurun_out/${TAG}_k_*.ncu-rep; do case $f in *k_odom_assoc.ncu-rep) ;; *) rm -f $f;; esac; done
python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench.json')); print(d['value'], d['ms_per_step'], d['e2e'], d.get('cpu_baseline'), {k:v for k,v in d['roofline'].items() if k!='kernels'})"
cat gpurun_out/${TAG}_bench_reference.json | head -c 600
