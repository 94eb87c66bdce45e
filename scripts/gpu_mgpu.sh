cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L | head -8
timeout 600 python -m pytest tests/test_gpu_mapping.py -m gpu -x -q -k "multi_process or two_contexts" 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29701 scripts/bench_config5.py --check > gpurun_out/config5_n$N.json 2> gpurun_out/config5_n$N.err
tail -3 gpurun_out/config5_n$N.json | cut -c1-1500; tail -5 gpurun_out/config5_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29702 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 400 gpurun_out/bench_n$N.json
