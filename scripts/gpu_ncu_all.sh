# launch list (gpu__time_duration) of a short steady-state run + ncu --set full of the top kernels
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${1:-r01b}
LL_B=128 LL_STEPS=10 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 168 -c 72 --csv --log-file gpurun_out/${TAG}_launches.csv python scripts/prof_run.py > gpurun_out/${TAG}_launches.log 2>&1
for k in k_ring_lessflat k_odom_assoc k_ring_sort k_classify k_lm_solve_odom k_odom_prep k_scatter k_ring_pick; do
LL_B=128 LL_STEPS=10 timeout 600 ncu --set full --clock-control none --import-source on -k regex:^$k\$ -s 8 -c 1 -o gpurun_out/${TAG}_$k -f python scripts/prof_run.py > gpurun_out/${TAG}_$k.log 2>&1
done
LL_B=128 LL_STEPS=10 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_odom_assoc_heavy -s 8 -c 1 -o gpurun_out/${TAG}_k_odom_assoc_heavy -f python scripts/prof_run.py > gpurun_out/${TAG}_heavy.log 2>&1
ls -la gpurun_out/*.ncu-rep | wc -l
