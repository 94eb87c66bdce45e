cd $GRAFT_REPO_ROOT
python scripts/time_latency.py
LL_GRAPH=0 python scripts/time_latency.py
