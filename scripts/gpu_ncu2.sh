# usage: bash scripts/gpu_ncu2.sh  — ncu --set full of the NN thread pass in both forms (256 lanes)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
LL_B=256 LL_STEPS=10 timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_odom_assoc$ -s 24 -c 1 -o gpurun_out/r02_assoc_global -f python scripts/prof_run.py > gpurun_out/r02_assoc_global.log 2>&1
tail -2 gpurun_out/r02_assoc_global.log
LL_ASSOC_SLAB=1 LL_B=256 LL_STEPS=10 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_odom_assoc_slab -s 24 -c 1 -o gpurun_out/r02_assoc_slab -f python scripts/prof_run.py > gpurun_out/r02_assoc_slab.log 2>&1
tail -2 gpurun_out/r02_assoc_slab.log
ls -la gpurun_out/*.ncu-rep
