cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for h in 1.3 1.15; do echo "H=$h"; LL_GRID_H=$h python scripts/prof_kernels.py assoc 2>&1 | tail -1; done
for hb in 296 1184; do echo "HB=$hb"; LL_HEAVY_BLOCKS=$hb python scripts/prof_kernels.py assoc 2>&1 | tail -1; done
python bench.py --no-cpu > gpurun_out/s4_bench.json; python -c "
import json; d=json.load(open('gpurun_out/s4_bench.json')); print(d['value'], d['ms_per_step'], d['e2e'], {k:v['ms_per_launch'] for k,v in list(d['roofline']['kernels'].items())[:8]})"
