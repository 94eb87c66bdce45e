cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/time_latency.py
LL_GRAPH=0 python scripts/time_latency.py
LL_GRAPH=0 LL_LM_PARTS=1 python scripts/time_latency.py
LL_B=1 LL_GRAPH=0 python scripts/prof_kernels.py lm_solve
