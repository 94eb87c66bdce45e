cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for t in 512 256 128; do echo "LM_THREADS=$t"; LL_B=256 LL_LM_THREADS=$t timeout 300 python scripts/prof_kernels.py lm_solve; done
echo "B=128"; for t in 512 256; do LL_B=128 LL_LM_THREADS=$t timeout 300 python scripts/prof_kernels.py lm_solve; done
