cd $GRAFT_REPO_ROOT
echo "W4 minb8"; LL_B=256 python scripts/prof_kernels.py assoc
echo "W8 minb8"; LL_LIB_PATH=$GRAFT_REPO_ROOT/build_w8/liblightloam_w8.so LL_B=256 python scripts/prof_kernels.py assoc
echo "W8 minb6"; LL_ASSOC_MINB=6 LL_LIB_PATH=$GRAFT_REPO_ROOT/build_w8/liblightloam_w8.so LL_B=256 python scripts/prof_kernels.py assoc
echo "W4 minb6"; LL_ASSOC_MINB=6 LL_B=256 python scripts/prof_kernels.py assoc
echo "W4 minb10"; LL_ASSOC_MINB=10 LL_B=256 python scripts/prof_kernels.py assoc
