"""bench.py --config 1 / 3 / 4 / 5: the other BASELINE.json configurations (configs[0], [2], [3], [4]) as measured runs.
Each prints ONE JSON line in bench.py's format (metric / value / unit / config.workload / e2e / roofline where a kernel
dominates / cpu_baseline where the CPU path is what is asked for).  `--config 2` (the headline) lives in bench.py.

  1  VLP-16 single scan, feature extraction + odometry: restated reference CPU path, ms per stage (+ the GPU latency beside it)
  3  HDL-64 scan-to-map against a 1e6 + 1e5 point map, graph vote off and on, L2 flushed between timed steps
  4  10k-scan HDL-64 stream cut into contiguous segments over the GPUs (and into lanes inside a GPU), one all-gather
     of the segment transforms + prefix product; deviation from the unsegmented chain measured on rank 0
  5  HDL-32 scan-to-map, map in N x-slabs (one per GPU), 28-double all-reduce inside the LM kernel vs ncclAllReduce
"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
CONFIGS = json.load(open(os.path.join(ROOT, "BASELINE.json")))["configs"]


def _ll():
    return importlib.import_module("light-loam_b200")


def _orc():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc_py
    orc_py.build()
    return orc_py


def _dist_init(world, local_rank):
    import torch
    sys.path.insert(0, ROOT)
    from bench import quiet_nccl
    torch.cuda.set_device(local_rank)
    if world <= 1:
        return None
    import torch.distributed as dist
    quiet_nccl()
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    return dist


def config3_map(n_s=1000000, n_c=100000, seed=3):
    """1e6 surf + 1e5 corner map points of a box room (ground, four walls, poles) in the 5 x 5 x 3 cubes."""
    rng = np.random.default_rng(seed)
    surf = np.zeros((n_s, 4), np.float32)
    u = rng.uniform(-1, 1, (n_s, 2))
    which = rng.integers(0, 5, n_s)
    surf[:, 0] = np.where(which == 1, 60, np.where(which == 2, -60, u[:, 0] * 60))
    surf[:, 1] = np.where(which == 3, 40, np.where(which == 4, -40, u[:, 1] * 40))
    surf[:, 2] = np.where(which == 0, -1.73, rng.uniform(-1.73, 13, n_s))
    surf[:, :3] += rng.normal(0, 0.01, (n_s, 3))
    corner = np.zeros((n_c, 4), np.float32)
    poles = rng.uniform(-55, 55, (200, 2))
    pid = rng.integers(0, 200, n_c)
    corner[:, 0], corner[:, 1] = poles[pid, 0], poles[pid, 1] * 0.7
    corner[:, 2] = rng.uniform(-1.7, 8, n_c)
    corner[:, :3] += rng.normal(0, 0.01, (n_c, 3))
    return corner, surf


# ---------------------------------------------------------------------------------------------------------------------
# config 1: VLP-16 on the reference CPU path
# ---------------------------------------------------------------------------------------------------------------------
def run_config1(args, rank, world, local_rank):
    if rank != 0:
        return
    ll, orc = _ll(), _orc()
    n = max(args.steps, 10)
    scans = [ll.synth.scan(16, k) for k in range(n + 3)]
    pipe = orc.Pipeline(orc.config(16), with_mapping=False)
    odo_cfg = orc.config(16)
    ms = []
    for k, s in enumerate(scans):
        r = pipe.step(s)
        if k >= 3:
            ms.append(r["ms"][:2].copy())
    ms = np.array(ms)
    # linearisations per scan: 3 Solves x (1 + accepted steps); measured from the oracle's own summaries on the same scans
    odo = orc.Odometry(odo_cfg)
    lin = []
    for k, s in enumerate(scans):
        f = orc.extract_features(s, odo_cfg)
        odo.step(f["sharp"], f["less_sharp"], f["flat"], f["less_flat"])
        if k >= 3:
            lin.append(sum(int(st[5]) for st in odo.stats()))
    cpu = {"ms_extract_features": round(float(ms[:, 0].mean()), 3), "ms_odometry_3_outer": round(float(ms[:, 1].mean()), 3),
           "linearisations_per_scan": round(float(np.mean(lin)), 2),
           "ms_per_gn_iteration": round(float(ms[:, 1].mean()) / max(float(np.mean(lin)), 1.0), 3),
           "value": round(1e3 / float(ms.sum(1).mean()), 2), "unit": "scans/s", "cores": 1, "kind": "port",
           "sample": "%d consecutive VLP-16 scans (16 x 1000 pts), restated reference CPU path (PCL/Ceres unavailable offline), steady_clock per stage" % n}
    gpu = None
    try:
        import torch
        if torch.cuda.is_available():
            ctx = ll.Context(scan_line=16, device=local_rank)
            for s in scans[:3]:
                ctx.process_scans([s])
            t0 = time.perf_counter()
            for s in scans[3:]:
                ctx.process_scans([s])
            gpu = (time.perf_counter() - t0) / n * 1e3
            launches = ctx.stats().kernel_launches
            ctx.close()
    except Exception:
        gpu = None
    line = {"metric": "ms per scan (VLP-16, feature extraction + scan-to-scan odometry)", "value": round(gpu, 4) if gpu else None, "unit": "ms", "n_gpus": 1,
            "steps": n, "warmup": 3, "ms_per_step": round(gpu, 4) if gpu else None, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (features, NN) + f64 (residuals, LM)", "data": "synthetic",
            "config": {"workload": CONFIGS[0], "note": "the configuration is the reference's own CPU-runnable case: cpu_baseline carries its ms per stage; value = the same scans through ll_process_scans on one B200, single stream, synchronous (host scan in, pose out)"},
            "e2e": {"value": round(gpu, 4) if gpu else None, "unit": "ms", "h2d_bytes_per_step": int(scans[0].nbytes), "d2h_bytes_per_step": 14 * 8 + 4},
            "gpu_launches": int(launches * n) if gpu else 0, "cpu_baseline": cpu}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# config 3: scan-to-map against a 1e6-point map
# ---------------------------------------------------------------------------------------------------------------------
def run_config3(args, rank, world, local_rank):
    if rank != 0:
        return
    import torch
    ll = _ll()
    torch.cuda.set_device(local_rank)
    sys.path.insert(0, ROOT)
    from bench import peaks
    corner, surf = config3_map()
    q0 = np.array([0, 0, np.sin(np.pi / 4), np.cos(np.pi / 4)])
    t0 = np.array([25.0, 0.0, 0.0])
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # > 126 MB of L2
    steps, W = max(args.steps, 3), max(args.warmup, 3)
    out = {}
    for vote in (0, 1):
        ctx = ll.Context(scan_line=64, map_capacity=1 << 21, map_graph_vote=vote, device=local_rank)
        f = ctx.extract_features(ll.synth.scan(64, 0, mode=1))
        wall, dev, kern, launches = [], [], {}, 0
        for r in range(W + steps):
            ctx.reset()
            ctx.map_insert(corner, surf)       # the step's own voxel filter (LM:2155-2168) thins the map: re-insert before every step
            flush.fill_(r & 0xFF)              # L2 flush between timed iterations
            torch.cuda.synchronize()
            ctx.profile_enable(True)
            t_0 = time.perf_counter()
            m = ctx.mapping_step(f["less_sharp"], f["less_flat"], q0, t0)
            dt = time.perf_counter() - t_0
            prof = ctx.profile_read()
            ctx.profile_enable(False)
            if r >= W:
                wall.append(dt)
                dev.append(sum(v[0] for v in prof.values()))
                launches += ctx.stats().kernel_launches
                for k, v in prof.items():
                    a = kern.setdefault(k, [0.0, 0])
                    a[0] += v[0]
                    a[1] += v[1]
        st = ctx.stats()
        out[vote] = dict(wall_ms=float(np.median(wall)) * 1e3, dev_ms=float(np.median(dev)), kern=kern, st=st, pose=m, launches=launches)
        ctx.close()
    peak, peak_src = peaks()
    base = out[1]
    st = base["st"]
    Q, M = st.stack_corner + st.stack_surf, st.map_corner + st.map_surf
    ka = base["kern"]["k_map_assoc"]
    assoc_ms = ka[0] / ka[1]
    alg = 16 * Q + 16 * M + 88 * Q       # queries + every local-map point once (upper bound 16 (Q + U), U <= M) + the dense records written
    tot = sum(v[0] for v in base["kern"].values())
    kernels = {k: {"ms_per_launch": round(v[0] / v[1], 4), "launches_per_step": v[1] / steps, "share": round(v[0] / tot, 4)}
               for k, v in sorted(base["kern"].items(), key=lambda kv: -kv[1][0])}
    line = {"metric": "scan-to-map steps/sec (HDL-64 features vs 1e6-point local map, 2 x <= 5 GN linearisations)", "value": round(1e3 / base["dev_ms"], 2), "unit": "steps/s",
            "n_gpus": 1, "steps": steps, "warmup": W, "ms_per_step": round(base["dev_ms"], 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (5-NN) + f64 (fits, residuals, LM)", "data": "synthetic",
            "config": {"workload": CONFIGS[2], "map_points": [int(st.map_corner), int(st.map_surf)], "stack_points": [int(st.stack_corner), int(st.stack_surf)],
                       "correspondences": [int(st.map_corner_corr), int(st.map_surf_corr)], "graph_vote": "on (LM:2057-2072 enabled: %d of %d planes selected and doubled)" % (st.map_vote_selected, st.map_vote_corr),
                       "l2": "flushed between timed steps (256 MB fill)", "timing": "value = sum of the step's kernel times (CUDA event pairs per launch); single stream"},
            "vote_off": {"ms_per_step": round(out[0]["dev_ms"], 4), "wall_ms": round(out[0]["wall_ms"], 4), "pose_t": [float(x) for x in out[0]["pose"]["t"]]},
            "vote_on": {"ms_per_step": round(out[1]["dev_ms"], 4), "wall_ms": round(out[1]["wall_ms"], 4), "pose_t": [float(x) for x in out[1]["pose"]["t"]]},
            "e2e": {"value": round(1e3 / base["wall_ms"], 2), "unit": "steps/s", "h2d_bytes_per_step": int(f["less_sharp"].nbytes + f["less_flat"].nbytes + 56),
                    "d2h_bytes_per_step": 1600, "api": "ll_mapping_step (host feature clouds in, pose out, synchronous)"},
            "gpu_launches": int(base["launches"]),
            "roofline": {"bound": "hbm", "kernel": "k_map_assoc", "achieved": round(alg / (assoc_ms * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                         "frac": round(alg / (assoc_ms * 1e-3) / 1e9 / peak, 4), "traffic": None, "peak_source": peak_src, "alg_bytes_per_launch": int(alg),
                         "alg_bytes_note": "16 Q + 16 M + 88 Q: upper bound of 16 (Q + U), every local-map point counted once", "ms_per_launch": round(assoc_ms, 4),
                         "share_of_step": round(ka[0] / tot, 4), "kernels": kernels}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
# config 4: one long stream in contiguous segments
# ---------------------------------------------------------------------------------------------------------------------
def _qmul(a, b):
    ax, ay, az, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
    bx, by, bz, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
    return np.stack([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz, aw * bz + az * bw + ax * by - ay * bx,
                     aw * bw - ax * bx - ay * by - az * bz], -1)


def _qrot(q, v):
    u, w = q[..., :3], q[..., 3:4]
    uv = 2.0 * np.cross(u, v)
    return v + w * uv + np.cross(u, uv)


def chain_prefix(ends):
    """ends: (K, 7) end transform of every sub-segment in stream order (its last scan's pose in the frame of its first).
    Returns (K, 7): pose of each sub-segment's first scan in the stream's frame (LO:830-831 chaining: q = q_a * q_b,
    t = t_a + q_a * t_b).  Composition is associative, so the prefix is a log-step (Hillis-Steele) scan of vectorised
    quaternion operations - 9 passes for 512 sub-segments instead of a Python loop."""
    K = len(ends)
    q, t = ends[:, :4].copy(), ends[:, 4:7].copy()          # inclusive scan in place
    d = 1
    while d < K:
        qa, ta = q[:-d], t[:-d]                               # element i - d (earlier) composed with element i
        tn = ta + _qrot(qa, t[d:])
        qn = _qmul(qa, q[d:])
        q = np.concatenate([q[:d], qn])
        t = np.concatenate([t[:d], tn])
        d *= 2
    out = np.zeros((K, 7))
    out[0, 3] = 1.0
    out[1:, :4], out[1:, 4:] = q[:-1], t[:-1]                 # exclusive: the pose at which sub-segment k starts
    return out


def _qinv_apply(P0, P):
    """inverse(P0) o P for (..., 7) pose arrays (unit quaternions up to rounding: the conjugate is the inverse)."""
    qi = P0[..., :4] * np.array([-1.0, -1.0, -1.0, 1.0])
    return np.concatenate([_qmul(qi, P[..., :4]), _qrot(qi, P[..., 4:7] - P0[..., 4:7])], -1)


def run_config4(args, rank, world, local_rank):
    import torch
    ll = _ll()
    sys.path.insert(0, ROOT)
    from bench import pin_to_gpu_numa
    pin_to_gpu_numa(local_rank)
    dist = _dist_init(world, local_rank)
    S, L, K = args.stream_scans, args.lanes, max(1, args.overlap)
    seg = ll.multigpu.segment_ranges(S, world)[rank]
    subs = [(seg[0] + a, seg[0] + b) for a, b in ll.multigpu.segment_ranges(seg[1] - seg[0], L)]      # this rank's sub-segments, stream order
    # Sub-segment l owns the poses of scans [b, e).  It starts K scans early (scan b - 1 is the anchor that ties it to its
    # predecessor; the K - 1 before that only re-converge the warm start and the graph-vote gate) and throws those poses away.
    first = [max(0, b - K) for b, _ in subs]
    anchor = [(b - 1 - f) if b > 0 else 0 for f, (b, _) in zip(first, subs)]                         # local index of the anchor scan
    nsteps = [e - f for f, (_, e) in zip(first, subs)]
    order = sorted(range(L), key=lambda l: -nsteps[l])                                               # slot -> sub-segment: the longest first (a prefix stays active)
    T = max(nsteps)
    g0 = min(first)
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
        scans = list(ex.map(lambda g: ll.synth.scan(64, g, mode=1, scan_id=g), range(g0, seg[1])))
    ctx = ll.Context(scan_line=64, batch=L, device=local_rank)
    ctx.pool_upload(scans)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def maxreduce(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n_act = [sum(1 for l in order if nsteps[l] > t) for t in range(T)]
    slot_of = np.argsort(np.array(order))                                                             # sub-segment -> slot
    last_idx = np.array([nsteps[l] - 1 for l in range(L)])
    anchor_idx = np.array(anchor)

    def run_stream(feed):
        """All sub-segments in lockstep; returns the poses in the stream's frame and the seconds spent in the exchange."""
        ctx.reset()     # every lane starts like a fresh stream: identity warm start, frame counter 0 (no graph vote on its first 5 pairs, LO:794)
        local = np.zeros((L, T, 7))
        local[:, :, 3] = 1.0
        launches = 0
        for t in range(T):
            poses = feed(t, n_act[t])
            launches += ctx.L.ll_launch_count(ctx.h)
            local[np.array(order[:n_act[t]]), t] = poses[:, :7]
        t_x = time.perf_counter()
        rel = _qinv_apply(local[np.arange(L), anchor_idx][:, None, :], local)                            # poses relative to each lane's anchor scan
        ends = torch.from_numpy(rel[np.arange(L), last_idx]).to("cuda")                                  # (L, 7): anchor -> last scan
        if dist is not None:
            allends = torch.empty((world * L, 7), dtype=torch.float64, device="cuda")
            dist.all_gather_into_tensor(allends, ends)                                                    # the one exchange: N x L x 56 bytes over NVLink
        else:
            allends = ends
        starts = chain_prefix(allends.cpu().numpy())[rank * L:(rank + 1) * L]                             # stream-frame pose of every anchor
        q0, t0 = starts[:, None, :4], starts[:, None, 4:]
        glob = np.concatenate([_qmul(q0, rel[..., :4]), t0 + _qrot(q0, rel[..., 4:7])], -1)               # (L, T, 7), valid from the anchor on
        return glob, time.perf_counter() - t_x, launches

    ids = np.array([[first[l] - g0 + min(t, nsteps[l] - 1) for l in order] for t in range(T)], np.int32)       # pool index per (step, slot)

    def feed_pool(t, n):
        return ctx.process_pool(ids[t, :n])

    run_stream(feed_pool)                                      # warm-up pass over the whole segment (allocations, NCCL communicator, clocks)
    barrier()
    t0 = time.perf_counter()
    glob, t_exchange, launches = run_stream(feed_pool)
    barrier()
    dt = maxreduce(time.perf_counter() - t0)
    value = S / dt

    # ---- e2e: the same with every scan crossing PCIe (step-major pinned arena of packed xyz, one copy per step) ----------
    n_pts = [len(s) for s in scans]
    offs = np.zeros((T, L + 1), np.int64)
    total = 0
    for t in range(T):
        for slot, l in enumerate(order):
            offs[t, slot] = total
            if nsteps[l] > t:
                total += n_pts[ids[t, slot]] * 12
        offs[t, L] = total
    arena_t = torch.empty(total, dtype=torch.uint8).pin_memory()
    arena = arena_t.numpy()
    cnts = np.zeros((T, L), np.int32)
    for t in range(T):
        for slot, l in enumerate(order):
            if nsteps[l] > t:
                j = ids[t, slot]
                cnts[t, slot] = n_pts[j]
                arena[offs[t, slot]:offs[t, slot] + n_pts[j] * 12] = np.ascontiguousarray(scans[j][:, :3]).view(np.uint8).reshape(-1)

    def feed_packed(t, n):
        """Step t's poses; step t + 1 is already on its way (two submissions in flight: its copy overlaps step t's kernels)."""
        if t == 0:
            ctx.submit_packed(arena, offs[0, :n], cnts[0, :n], 12)
        if t + 1 < T:
            ctx.submit_packed(arena, offs[t + 1, :n_act[t + 1]], cnts[t + 1, :n_act[t + 1]], 12)
        return ctx.collect()

    run_stream(feed_packed)            # warm-up of the asynchronous path
    barrier()
    t0 = time.perf_counter()
    run_stream(feed_packed)
    barrier()
    e2e_dt = maxreduce(time.perf_counter() - t0)

    # ---- deviation from the unsegmented chain (rank 0, its own segment): what cutting the stream costs -----------------
    dev = None
    if rank == 0:
        one = ll.Context(scan_line=64, batch=1, device=local_rank)
        chain = np.zeros((len(scans), 7))
        for k, s in enumerate(scans):
            chain[k] = one.process_scans([s])[0][:7]
        one.close()
        mine = np.zeros((len(scans), 7))
        for l in range(L):
            b, e = subs[l]
            mine[b - g0:e - g0] = glob[l, b - first[l]:e - first[l]]      # rank 0: the stream's frame = its segment's frame
        dtm = np.abs(mine[:, 4:] - chain[:, 4:]).max(1)
        ang = 2 * np.arccos(np.minimum(1.0, np.abs((mine[:, :4] * chain[:, :4]).sum(1))))
        # per-scan increments in the sensor frame (inverse(pose k-1) * pose k): what a scan pair's solve changes when the lane
        # restarted K scans before

        def increments(P):
            r = _qinv_apply(P[:-1], P[1:])
            return r[:, 4:], r[:, :4]
        (ti_m, qi_m), (ti_c, qi_c) = increments(mine), increments(chain)
        inc_t = np.linalg.norm(ti_m - ti_c, axis=1)
        inc_r = 2 * np.arccos(np.minimum(1.0, np.abs((qi_m * qi_c).sum(1))))
        dev = {"scans": len(scans), "sub_segments": L, "overlap_scans": K, "max_abs_translation_m": float(dtm.max()), "max_rotation_rad": float(ang.max()),
               "end_of_segment_translation_m": float(dtm[-1]), "max_per_scan_increment_m": float(inc_t.max()), "median_per_scan_increment_m": float(np.median(inc_t)),
               "max_per_scan_increment_rad": float(inc_r.max()), "increments_within_1e-4": float(np.mean((inc_t < 1e-4) & (inc_r < 1e-4))), "path_length_m": float(len(scans) - 1)}
    if rank == 0:
        line = {"metric": "scans/sec (HDL-64, 130k pts, 5 GN iters)", "value": round(value, 1), "unit": "scans/s", "n_gpus": world, "steps": T, "warmup": T,
                "ms_per_step": round(dt / T * 1e3, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32 (features, NN) + f64 (residuals, LM)", "data": "synthetic",
                "config": {"workload": CONFIGS[3], "stream_scans": S, "segments": world, "sub_segments_per_gpu": L, "overlap_scans": K,
                           "steps_per_gpu": T, "scans_processed_incl_overlap": int(sum(nsteps)) * world,
                           "parallelism": "contiguous segments per GPU, %d lanes per GPU, each starting %d scans early; one ncclAllGather of the %d x 7-double segment transforms + log-step prefix product" % (L, K, world * L),
                           "l2": "inputs larger than L2: %d lanes x 2.08 MB per step" % L},
                "exchange": {"collective": "all_gather_into_tensor (NCCL, device tensors) of %d x 56 B + prefix product + pose placement" % (world * L), "ms": round(t_exchange * 1e3, 3),
                             "share_of_run": round(t_exchange / dt, 5)},
                "e2e": {"value": round(S / e2e_dt, 1), "unit": "scans/s", "h2d_bytes_per_step": int(total / T), "d2h_bytes_per_step": L * (14 * 8 + 4),
                        "api": "ll_submit_packed / ll_collect per step (step-major pinned arena of packed xyz), poses read back every step"},
                "gpu_launches": int(launches), "deviation_vs_unsegmented_chain": dev}
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------------
# config 5: map slabs + all-reduce of the normal equations
# ---------------------------------------------------------------------------------------------------------------------
def run_config5(args, rank, world, local_rank):
    import torch
    ll = _ll()
    dist = _dist_init(world, local_rank)
    n_s, n_c = (300000, 30000) if args.small else (1000000, 100000)
    corner, surf = config3_map(n_s=n_s, n_c=n_c, seed=7)
    ctx = ll.Context(scan_line=32, map_capacity=1 << 21, device=local_rank)
    if dist is not None:
        lo, hi = ll.multigpu.attach_all(ctx, dist, -60.0, 60.0)
        mine_c = np.ascontiguousarray(ll.multigpu.slab_with_halo(corner, lo, hi))
        mine_s = np.ascontiguousarray(ll.multigpu.slab_with_halo(surf, lo, hi))
    else:
        mine_c, mine_s = corner, surf
    f = ctx.extract_features(ll.synth.scan(32, 0, mode=1))
    q0 = np.array([0, 0, np.sin(np.pi / 4), np.cos(np.pi / 4)])
    t0 = np.array([25.0, 0.0, 0.0])
    steps, W = max(args.steps, 3), max(args.warmup, 3)
    times, kern, m, launches = [], {}, None, 0
    for r in range(W + steps):
        ctx.reset()
        ctx.map_insert(mine_c, mine_s)
        ctx.profile_enable(r >= W)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        t_0 = time.perf_counter()
        m = ctx.mapping_step(f["less_sharp"], f["less_flat"], q0, t0)
        dt = torch.tensor([time.perf_counter() - t_0], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if r >= W:
            times.append(float(dt.item()))
            launches += ctx.stats().kernel_launches
            for k, v in ctx.profile_read().items():
                a = kern.setdefault(k, [0.0, 0])
                a[0] += v[0]
                a[1] += v[1]
    st = ctx.stats()
    # the baseline the in-kernel all-reduce replaces: one ncclAllReduce of the 28 doubles per evaluation (<= 2 x (5 + 4) per step)
    nccl_us = None
    if dist is not None:
        buf = torch.zeros(28, dtype=torch.float64, device="cuda")
        for _ in range(20):
            dist.all_reduce(buf)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            dist.all_reduce(buf)
        e1.record()
        torch.cuda.synchronize()
        nccl_us = e0.elapsed_time(e1) / 200 * 1e3
    poses_same, single = True, None
    if dist is not None:
        poses = [torch.zeros(7, dtype=torch.float64, device="cuda") for _ in range(world)]
        dist.all_gather(poses, torch.tensor(np.concatenate([m["q"], m["t"]]), device="cuda"))
        poses_same = all(torch.equal(poses[0], p) for p in poses)
        corr = torch.tensor([st.map_corner_corr, st.map_surf_corr], dtype=torch.int64, device="cuda")
        dist.all_reduce(corr)
        if rank == 0:
            ref = ll.Context(scan_line=32, map_capacity=1 << 21, device=local_rank)
            ref.map_insert(corner, surf)
            w = ref.mapping_step(f["less_sharp"], f["less_flat"], q0, t0)
            s1 = ref.stats()
            single = {"max_abs_dq": float(np.abs(w["q"] - m["q"]).max()), "max_abs_dt": float(np.abs(w["t"] - m["t"]).max()),
                      "corr_single": [s1.map_corner_corr, s1.map_surf_corr], "corr_sum_over_ranks": corr.tolist()}
            ref.close()
        dist.barrier()
    if rank == 0:
        evals = int(sum(st.map_jacobian_evals)) + 8          # linearisations + up to 4 cost-only evaluations per Solve
        solve = kern.get("k_lm_solve_map", [0.0, 1])
        ms = float(np.median(times)) * 1e3
        line = {"metric": "ms per scan-to-map step (HDL-32, map in x-slabs, all-reduce of the 6x6 normal equations)", "value": round(ms, 4), "unit": "ms", "n_gpus": world,
                "steps": steps, "warmup": W, "ms_per_step": round(ms, 4), "higher_is_better": False, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32 (5-NN) + f64 (fits, residuals, LM, all-reduce)", "data": "synthetic",
                "config": {"workload": CONFIGS[4], "slabs": world, "map_points": [n_c, n_s], "halo_m": 1.0,
                           "parallelism": "one x-slab (+ 1 m halo) of the map per GPU; every rank runs the same ll_mapping_step; the LM kernel all-reduces 28 doubles per evaluation over peer-memory mailboxes (no NCCL call on the data path)"},
                "allreduce": {"in_kernel": {"ms_per_solve_kernel": round(solve[0] / max(solve[1], 1), 4), "evaluations_per_step_max": evals,
                                            "poses_identical_across_ranks": bool(poses_same)},
                              "nccl_baseline": {"us_per_allreduce_28_doubles": round(nccl_us, 2) if nccl_us else None,
                                                "ms_per_step_if_every_evaluation_called_it": round(nccl_us * evals / 1e3, 4) if nccl_us else None,
                                                "note": "latency of one ncclAllReduce of 224 bytes on the same GPUs (torch.distributed, CUDA events, 200 calls); a host-driven loop pays it plus a launch per evaluation"}},
                "vs_single_gpu": single,
                "e2e": {"value": round(ms, 4), "unit": "ms", "h2d_bytes_per_step": int(f["less_sharp"].nbytes + f["less_flat"].nbytes + 56), "d2h_bytes_per_step": 1600,
                        "api": "ll_mapping_step (host feature clouds in, pose out), max over ranks"},
                "gpu_launches": int(launches),
                "kernels_rank0_ms_per_launch": {k: round(v[0] / max(v[1], 1), 4) for k, v in sorted(kern.items(), key=lambda kv: -kv[1][0])[:8]}}
        print(json.dumps(line))
        if args.check:
            ok = poses_same and (single is None or (single["max_abs_dq"] < 1e-9 and single["max_abs_dt"] < 1e-9 and single["corr_single"] == single["corr_sum_over_ranks"]))
            print("config5 ok" if ok else "config5 MISMATCH")
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
