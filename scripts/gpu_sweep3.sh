cd $GRAFT_REPO_ROOT
export LL_B=256
echo base; timeout 300 python scripts/prof_kernels.py assoc index_s
echo "AZ=128/512 K4 D16"; LL_AZ_CORNER=128 LL_AZ_SURF=512 LL_ASSOC_KMAX=4 LL_ASSOC_DMAX=16 timeout 300 python scripts/prof_kernels.py assoc index_s
echo "AZ=64/512 K4 D16"; LL_AZ_CORNER=64 LL_AZ_SURF=512 LL_ASSOC_KMAX=4 LL_ASSOC_DMAX=16 timeout 300 python scripts/prof_kernels.py assoc index_s
echo "AZ=128/512 K3 D12"; LL_AZ_CORNER=128 LL_AZ_SURF=512 LL_ASSOC_KMAX=3 LL_ASSOC_DMAX=12 timeout 300 python scripts/prof_kernels.py assoc index_s
echo "AZ=128/256 K2 D8"; LL_AZ_CORNER=128 LL_AZ_SURF=256 timeout 300 python scripts/prof_kernels.py assoc index_s
echo "MINB=6"; LL_ASSOC_MINB=6 timeout 300 python scripts/prof_kernels.py assoc
echo "MINB=10"; LL_ASSOC_MINB=10 timeout 300 python scripts/prof_kernels.py assoc
