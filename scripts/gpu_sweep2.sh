cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
LL_B=256 timeout 300 python scripts/prof_kernels.py
for m in 5 6 8; do echo "CLS_MINB=$m"; LL_B=256 LL_CLS_MINB=$m timeout 300 python scripts/prof_kernels.py classify; done
for m in 4 6 8; do echo "SCT_MINB=$m"; LL_B=256 LL_SCT_MINB=$m timeout 300 python scripts/prof_kernels.py scatter; done
