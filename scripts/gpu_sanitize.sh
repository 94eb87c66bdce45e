# compute-sanitizer memcheck + racecheck over short runs of every kernel family (scripts/sanitize_run.py)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${1:-r02}
for tool in memcheck racecheck; do
# few lanes: warp-per-query association, cluster solves, graph replay
LL_STEPS=8 timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_run.py default > gpurun_out/${TAG}_sanitize_${tool}_default.log 2>&1; tail -3 gpurun_out/${TAG}_sanitize_${tool}_default.log
# the forms the batched path uses: thread pass + queue, one CTA per solve
LL_ASSOC_DIRECT=0 LL_LM_CLUSTER=1 LL_VOTE_FUSED=0 LL_STEPS=8 timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_run.py default > gpurun_out/${TAG}_sanitize_${tool}_batched.log 2>&1; tail -3 gpurun_out/${TAG}_sanitize_${tool}_batched.log
LL_ASSOC_SLAB=1 LL_STEPS=4 timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_run.py default > gpurun_out/${TAG}_sanitize_${tool}_slab.log 2>&1; tail -3 gpurun_out/${TAG}_sanitize_${tool}_slab.log
LL_STEPS=8 timeout 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_run.py modes > gpurun_out/${TAG}_sanitize_${tool}_modes.log 2>&1; tail -3 gpurun_out/${TAG}_sanitize_${tool}_modes.log
done
