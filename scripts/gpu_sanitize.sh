# compute-sanitizer memcheck + racecheck over a short fused-pipeline run (2 lanes, 64 lines)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
LL_B=2 LL_STEPS=4 timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python scripts/prof_run.py > gpurun_out/sanitize_memcheck.log 2>&1; tail -4 gpurun_out/sanitize_memcheck.log
LL_B=2 LL_STEPS=3 timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python scripts/prof_run.py > gpurun_out/sanitize_racecheck.log 2>&1; tail -4 gpurun_out/sanitize_racecheck.log
