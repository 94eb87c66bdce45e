# usage (gpurun --gpus 8): bash scripts/gpu_mgpu8.sh
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${N:-8}
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
nproc > gpurun_out/r02_nproc_n$N.txt; free -g >> gpurun_out/r02_nproc_n$N.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > gpurun_out/r02_bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -2 gpurun_out/bench_n$N.err
timeout 600 $TR bench.py --gpus $N --config 5 --steps 6 --check > gpurun_out/r02_config5_n$N.json 2> gpurun_out/cfg5_n$N.err; tail -2 gpurun_out/cfg5_n$N.err
timeout 900 $TR bench.py --gpus $N --config 4 --stream-scans ${S:-10000} --lanes ${L:-64} --overlap 6 > gpurun_out/r02_config4_n$N.json 2> gpurun_out/cfg4_n$N.err; tail -2 gpurun_out/cfg4_n$N.err
python - <<PY
import json
for f in ("r02_bench_n$N","r02_config5_n$N","r02_config4_n$N"):
    try:
        for ln in open("gpurun_out/%s.json"%f):
            if ln.startswith("{"):
                d=json.loads(ln); print(f, d["n_gpus"], d["value"], d.get("e2e",{}).get("value"), d.get("e2e",{}).get("h2d_gbs"), d.get("exchange"), d.get("deviation_vs_unsegmented_chain"), d.get("allreduce"), d.get("vs_single_gpu"))
    except Exception as e: print(f, "ERR", e)
PY
