cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -k "vote_partial or deskew_fused" 2>&1 | tail -40
timeout 600 python scripts/exp_two_ctx.py 2>&1 | tail -4
