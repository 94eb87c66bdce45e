"""BASELINE.json configs[4]: HDL-32 scan-to-map with the map tiled into spatial slabs, one slab (+ 1 m halo) per GPU,
per-GPU JtJ, all-reduce of the 6x6 normal equations inside the LM kernel (peer-memory mailboxes, no NCCL call on the
data path).  Launch:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_config5.py
Prints one JSON line on rank 0.  --check additionally compares against a single-GPU context on rank 0."""
import argparse, importlib, importlib.util, json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ll = importlib.import_module("light-loam_b200")
spec = importlib.util.spec_from_file_location("prof_map", os.path.join(ROOT, "scripts", "prof_map.py"))
pm = importlib.util.module_from_spec(spec)
spec.loader.exec_module(pm)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=6)
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--replicated", action="store_true", help="every rank holds the whole map (default: slab + 1 m halo)")
    a = ap.parse_args()
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    n_s, n_c = (300000, 30000) if a.small else (1000000, 100000)
    corner, surf = pm.config3_map(n_s=n_s, n_c=n_c, seed=7)
    ctx = ll.Context(scan_line=32, map_capacity=1 << 21, device=lr)
    lo, hi = ll.multigpu.attach_all(ctx, dist, -60.0, 60.0)
    mine_c = corner if a.replicated else np.ascontiguousarray(ll.multigpu.slab_with_halo(corner, lo, hi))
    mine_s = surf if a.replicated else np.ascontiguousarray(ll.multigpu.slab_with_halo(surf, lo, hi))
    f = ctx.extract_features(ll.synth.scan(32, 0, mode=1))
    q0 = np.array([0, 0, np.sin(np.pi / 4), np.cos(np.pi / 4)])
    t0 = np.array([25.0, 0.0, 0.0])
    times, prof, m = [], None, None
    for r in range(a.reps):
        ctx.reset()
        ctx.map_insert(mine_c, mine_s)
        if r == a.reps - 1:
            ctx.profile_enable(True)
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t_0 = time.perf_counter()
        m = ctx.mapping_step(f["less_sharp"], f["less_flat"], q0, t0)
        dt = torch.tensor([time.perf_counter() - t_0], dtype=torch.float64, device="cuda")
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        times.append(float(dt.item()))
    prof = ctx.profile_read()
    st = ctx.stats()
    corr = torch.tensor([st.map_corner_corr, st.map_surf_corr, st.map_corner, st.map_surf], dtype=torch.int64, device="cuda")
    allc = [torch.zeros_like(corr) for _ in range(world)]
    dist.all_gather(allc, corr)
    poses = [torch.zeros(7, dtype=torch.float64, device="cuda") for _ in range(world)]
    dist.all_gather(poses, torch.tensor(np.concatenate([m["q"], m["t"]]), device="cuda"))
    same = all(torch.equal(poses[0], p) for p in poses)
    ok = True
    single = None
    if a.check and rank == 0:
        ref = ll.Context(scan_line=32, map_capacity=1 << 21, device=lr)
        ref.map_insert(corner, surf)
        w = ref.mapping_step(f["less_sharp"], f["less_flat"], q0, t0)
        s1 = ref.stats()
        dq, dtt = float(np.abs(w["q"] - m["q"]).max()), float(np.abs(w["t"] - m["t"]).max())
        tot = torch.stack(allc).sum(0).tolist()
        single = {"dq": dq, "dt": dtt, "corr_single": [s1.map_corner_corr, s1.map_surf_corr], "corr_sum": tot[:2]}
        ok = same and dq < 1e-9 and dtt < 1e-9 and tot[0] == s1.map_corner_corr and tot[1] == s1.map_surf_corr and m["rc"] == 0
        ref.close()
    dist.barrier()
    if rank == 0:
        kern = {k: round(v[0] / max(v[1], 1), 4) for k, v in prof.items() if k in ("k_lm_solve_map", "k_map_assoc", "k_grid_scatter", "k_map_gather")}
        line = {"config": "BASELINE.json configs[4]: HDL-32 scan-to-map, map in %d x-slabs (%s), per-GPU JtJ + in-kernel all-reduce of 28 doubles" % (world, "replicated map" if a.replicated else "slab + 1 m halo per GPU"),
                "n_gpus": world, "map_points": [n_c, n_s], "ms_per_step_median": round(float(np.median(times[1:])) * 1e3, 3), "ms_per_step_all": [round(t * 1e3, 3) for t in times],
                "kernel_ms_per_launch_rank0": kern, "jacobian_evals": list(st.map_jacobian_evals), "per_rank_corr_and_map": [c.tolist() for c in allc],
                "poses_identical_across_ranks": same, "vs_single_gpu": single, "pose": [float(x) for x in poses[0].tolist()]}
        print(json.dumps(line))
        if a.check:
            print("config5 ok" if ok else "config5 MISMATCH")
    ctx.close()
    dist.destroy_process_group()
    if a.check and rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
