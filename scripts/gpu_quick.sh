# parity tests + per-kernel device times of the fused pipeline (one short GPU call)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -${TAILN:-6}
for B in ${BS:-128}; do LL_B=$B timeout 300 python scripts/prof_kernels.py; done
