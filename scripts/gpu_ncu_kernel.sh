# usage: bash scripts/gpu_ncu_kernel.sh <kernel regex> <out name> [skip] [count]
set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
LL_B=128 LL_STEPS=10 timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s ${3:-24} -c ${4:-1} -o gpurun_out/$2 -f python scripts/prof_run.py > gpurun_out/$2.log 2>&1
tail -3 gpurun_out/$2.log
