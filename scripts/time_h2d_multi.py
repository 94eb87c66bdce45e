"""Raw pinned host-to-device bandwidth with all ranks copying at once (no kernels, no library): the link the end-to-end
figure is bound by.  torchrun --nproc-per-node N scripts/time_h2d_multi.py  -> one JSON line on rank 0."""
import json, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
bench.pin_to_gpu_numa(lr)
bench.quiet_nccl()
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
nbytes = 400 * 1000 * 1000
host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
for _ in range(3):
    dev.copy_(host, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
reps = 40
t0 = time.perf_counter()
for _ in range(reps):
    dev.copy_(host, non_blocking=True)
torch.cuda.synchronize()
gbs = nbytes * reps / (time.perf_counter() - t0) / 1e9
t = torch.tensor([gbs], dtype=torch.float64, device="cuda")
allg = [torch.zeros_like(t) for _ in range(world)]
if world > 1:
    dist.all_gather(allg, t)
else:
    allg = [t]
if rank == 0:
    per = [round(float(x.item()), 2) for x in allg]
    print(json.dumps({"n_gpus": world, "h2d_gbs_per_gpu": per, "aggregate_gbs": round(sum(per), 1), "copy": "40 x 400 MB pinned cudaMemcpyAsync per rank, all ranks at once"}))
if world > 1:
    dist.destroy_process_group()
