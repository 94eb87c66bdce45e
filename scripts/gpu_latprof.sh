# ncu --set full of the single-stream forms of the kernels (one lane, eager launches), summaries only
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
LL_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 270 -c 30 --csv --log-file gpurun_out/lat1_launches.csv python scripts/lat_run.py > /dev/null 2>&1
for k in k_odom_assoc_direct k_lm_solve_odom k_odom_prep_vote; do
LL_GRAPH=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:^$k\$ -s 12 -c 1 -o gpurun_out/lat1_$k -f python scripts/lat_run.py > gpurun_out/lat1_$k.log 2>&1
done
python scripts/ncu_summary.py gpurun_out/lat1_launches.csv gpurun_out/lat1_k_*.ncu-rep > gpurun_out/r02c_ncu_summary_lane1.json 2> gpurun_out/lat1_summary.err
for k in k_odom_assoc_direct k_lm_solve_odom; do python scripts/ncu_lines.py gpurun_out/lat1_$k.ncu-rep 20 > gpurun_out/r02c_lane1_${k}_lines.txt 2>/dev/null; done
rm -f gpurun_out/lat1_k_*.ncu-rep
python -c "
import json; s=json.load(open('gpurun_out/r02c_ncu_summary_lane1.json'))
for e in s['full']: print(e['kernel'][:60], e['gpu__time_duration.sum'], e['launch__grid_size'], e['launch__block_size'], e['sm__warps_active.avg.pct_of_peak_sustained_active'], e['smsp__issue_active.avg.pct_of_peak_sustained_active'])"
