cd $GRAFT_REPO_ROOT
LL_B=1 python scripts/prof_kernels.py
python scripts/time_latency.py
