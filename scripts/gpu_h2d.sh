cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python scripts/time_h2d.py 2>&1 | tee gpurun_out/time_h2d.txt
LL_PIN=0 timeout 600 python scripts/time_h2d.py 2>&1 | tail -4
