cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q -k "features" 2>&1 | tail -2
for v in "5 8" "6 8" "8 8" "8 6" "5 6"; do set -- $v; echo "SCT_MINB=$1 CLS_MINB=$2"; LL_SCT_MINB=$1 LL_CLS_MINB=$2 LL_B=256 python scripts/prof_kernels.py classify scatter lessflat; done
