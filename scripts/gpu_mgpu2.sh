# usage (gpurun --gpus N): N=<gpus> bash scripts/gpu_mgpu2.sh
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${N:-2}
nvidia-smi topo -m > gpurun_out/r02_topo_n$N.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q -k "multi_process or two_contexts or radix_sort or split_solve" 2>&1 | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
timeout 900 $TR bench.py --gpus $N --config 4 --stream-scans ${S:-2560} --lanes ${L:-128} > gpurun_out/r02_config4_n$N.json 2> gpurun_out/cfg4_n$N.err; tail -3 gpurun_out/cfg4_n$N.err; cut -c1-2600 gpurun_out/r02_config4_n$N.json
timeout 900 $TR bench.py --gpus $N --config 5 --steps 6 --check > gpurun_out/r02_config5_n$N.json 2> gpurun_out/cfg5_n$N.err; tail -3 gpurun_out/cfg5_n$N.err; cut -c1-2600 gpurun_out/r02_config5_n$N.json
timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu > gpurun_out/r02_bench_n$N.json 2> gpurun_out/bench_n$N.err; tail -3 gpurun_out/bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n$N.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e'], d['config'].get('cpu_affinity_cores'), d.get('parity',{}).get('status'))"
