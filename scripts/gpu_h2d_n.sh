# usage (gpurun --gpus N): N=<gpus> [BENCH=1] bash scripts/gpu_h2d_n.sh  — raw pinned H2D with all ranks copying at once (+ the bench at that N)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${N:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621"
timeout 300 $TR scripts/time_h2d_multi.py 2>/dev/null | grep '^{' > gpurun_out/r02_h2d_raw_n$N.json; cat gpurun_out/r02_h2d_raw_n$N.json
if [ -n "$BENCH" ]; then
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu 2> gpurun_out/bench_n$N.err | grep '^{' > gpurun_out/r02_bench_n$N.json; tail -2 gpurun_out/bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n$N.json')); print(d['n_gpus'], d['value'], d['e2e']['value'], d['e2e']['h2d_gbs'], d['parity']['status'])"
fi
