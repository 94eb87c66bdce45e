# parity tests, then association tuning sweeps (per-kernel device times)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python scripts/prof_kernels.py
for d in 3 5 12; do echo "DMAX=$d"; LL_ASSOC_DMAX=$d timeout 300 python scripts/prof_kernels.py assoc; done
for a in "128 512" "64 512"; do set -- $a; echo "AZ=$1/$2"; LL_AZ_CORNER=$1 LL_AZ_SURF=$2 timeout 300 python scripts/prof_kernels.py assoc index; done
LL_DEV_SKIP=16 LL_DEBUG_ASSOC=1 LL_B=128 LL_STEPS=10 timeout 300 python scripts/prof_run.py 2>&1 | tail -2
