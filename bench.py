#!/usr/bin/env python
"""bench.py — scans/sec of the Light-LOAM per-scan hot path on B200.

A "step" = one pass of the hot path (feature extraction SR:100-377 + scan-to-scan odometry LO:425-896:
3 outer iterations x <= 5 LM linearisations) over one batch of `--batch` independent HDL-64 scan streams
("lanes"), one new ~130k-point scan per lane per step.  Workload = BASELINE.json configs[1].

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA, through the C ABI)
  python bench.py --impl reference --gpus N ...            the reference's CPU algorithm (oracle restatement;
                                                           PCL/Ceres/ROS are not installable here) on the host cores

Prints ONE JSON line on rank 0.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "scans/sec (HDL-64, 130k pts, 5 GN iters)"
LAP = 157                 # one lap of the 25 m loop path at 1 m per scan
POOL_SCANS = 2 * LAP      # two laps with independent range noise (noise seed = pool index): a cyclic sequence of 314 distinct scans >= lanes
PARITY_LANES = 4          # lanes checked against the oracle after the timed loops
PARITY_STEPS = 9          # past now_frame > 5, so the graph vote is on for the last steps
NN_KERNEL = "k_odom_assoc"
WORKLOAD = "HDL-64 single scan (~130k pts), 5 GN iters, scan-to-scan odometry on 1xB200 (BASELINE.json configs[1]), batched over independent scan streams"


def quiet_nccl():
    """stdout carries ONE JSON line: NCCL's version banner (NCCL_DEBUG=VERSION / WARN, which some boxes export) must not
    join it.  An explicit INFO / TRACE request is left alone."""
    if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
        os.environ["NCCL_DEBUG"] = "NONE"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line): NVML polled every few
    milliseconds (nvidia_ml_py), nvidia-smi as the fallback."""

    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.sm, self.mx, self.seen = [], [], set()
        self.stop_flag = False
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll_nvml(self):
        n = self.nvml
        self.sm.append(float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)))
        self.mx.append(float(n.nvmlDeviceGetMaxClockInfo(self.h, n.NVML_CLOCK_SM)))
        get = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        r = int(get(self.h))
        bits = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        for k, b in bits.items():
            if r & b:
                self.seen.add(k)

    def _poll_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        r = [c.strip() for c in out.split(",")] if out else []
        if len(r) >= 6 and r[0].replace(".", "").isdigit():
            self.sm.append(float(r[0]))
            self.mx.append(float(r[1]))
            for i in range(4):
                if r[2 + i].lower().startswith("active"):
                    self.seen.add(self.NAMES[i])

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._poll_nvml()
                else:
                    self._poll_smi()
            except Exception:
                pass
            time.sleep(0.004 if self.nvml is not None else 0.1)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.seen), "samples": len(self.sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def make_pool(ll, n):
    """Pool scan j: pose j mod LAP of the loop path, range noise seeded by j (distinct scans across laps)."""
    from concurrent.futures import ThreadPoolExecutor   # the generator is a ctypes call: it releases the GIL
    with ThreadPoolExecutor(max_workers=min(32, os.cpu_count() or 1)) as ex:
        return list(ex.map(lambda j: ll.synth.scan(64, j % LAP, mode=1, scan_id=j), range(n)))


def lane_base(lanes, rank):
    return (rank * lanes * 5) % POOL_SCANS


def lane_ids(step, lanes, rank):
    """Lane i walks the cyclic scan sequence one scan ahead of lane i - 1: independent streams doing the same work, and the
    scans of one step are `lanes` consecutive pool entries (one contiguous block of the host arena for the e2e copy)."""
    return ((lane_base(lanes, rank) + step + np.arange(lanes)) % POOL_SCANS).astype(np.int32)


def pin_to_gpu_numa(index):
    """CPU affinity of this rank := the cores NVML reports as local to the GPU, BEFORE the pinned arena is allocated
    (first touch puts it on the GPU's NUMA node).  Returns the number of cores, or None when NVML cannot tell."""
    try:
        import pynvml
        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = int(vis.split(",")[index]) if vis and all(v.strip().isdigit() for v in vis.split(",")) else index
        h = pynvml.nvmlDeviceGetHandleByIndex(phys)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {w * 64 + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# per-scan algorithmic (compulsory) bytes of each kernel, from the measured point counts
def algorithmic_bytes(counts):
    Nr, P, nlf, nls, ns, nf = counts["n_raw"], counts["n_full"], counts["n_less_flat"], counts["n_less_sharp"], counts["n_sharp"], counts["n_flat"]
    feats = 16 * (ns + nls + nf) + 4 * (ns + nls + nf)
    return {
        "k_classify": 16 * Nr + 6 * Nr + 4 * 64 * (Nr / 256.0),   # read raw, write ring id (1 B) + azimuth (4 B) + rank (1 B) + tile histogram
        "k_scatter": 16 * Nr + 6 * Nr + 16 * P,              # read raw + ring/ori/rank, write ring-sorted cloud
        "k_ring_sort": 16 * P + 4 * P + 1 * P + 2 * P,         # read ring slabs (TMA), write curvature, label, sorted order
        "k_ring_pick": 2 * P + feats,                          # read sorted order (+ a few points), write picks
        "k_ring_lessflat": 2 * 16 * P + 1 * P + 16 * nlf,      # two passes over the ring (bbox, keys), labels, DS output
        "k_compact": 2 * (16 * nlf) + 2 * feats,
        "k_odom_assoc": 16 * (ns + nf) + 16 * (nls + nlf) + 8 * ns + 16 * nf,  # NN + ring window: queries + every indexed target once (upper bound 16 (Q + M)) + results
        "k_index_count": 16 * (nls + nlf),
        "k_index_scatter": 16 * (nls + nlf) + 16 * (nls + nlf),       # read the *Last clouds, write the bucket-ordered copy
        "k_lm_solve_odom": 88 * (ns + nf),                            # residual-block records, read once per solve (upper bound: every feature matched)
    }


def config_dict(B, world, n_raw):
    """The `config` object: identical in both arms (ours and --impl reference), so the two lines describe one workload."""
    return {"workload": WORKLOAD, "lanes_per_gpu": B, "scans_per_step": B * world, "points_per_scan": int(n_raw), "distinct_scans_in_pool": POOL_SCANS,
            "gn_linearisations_per_solve_max": 5, "outer_iterations": 3,
            "parallelism": "independent scan streams sharded per GPU, no data-path collective",
            "l2": "inputs larger than L2: %d lanes x %.2f MB raw scan = %.0f MB read per step (> 126 MB), no flush" % (B, n_raw * 16 / 1e6, B * n_raw * 16 / 1e6)}


def parity_check(ll, ctx, pool, B, rank):
    """Parity gate of THIS configuration in THIS run: the state is reset, PARITY_STEPS steps run through the pool path at
    the full lane count, and PARITY_LANES lanes (first, two inside, last) are compared with the CPU oracle fed the same
    scans: per-step pose (<= 1e-9 against the oracle in the device's voxel order, <= 1e-4 m / rad against the
    reference-faithful order) and, on the last step, the sharp / less-sharp / flat index lists (bit-exact)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc_py
    lanes = sorted({0, B // 3, (2 * B) // 3, B - 1})[:PARITY_LANES]
    ctx.reset()
    gpu = []
    for k in range(PARITY_STEPS):
        gpu.append(ctx.process_pool(lane_ids(k, B, rank))[lanes].copy())
    feats = {i: ctx.debug_features(i) for i in lanes}
    worst_exact, worst_faithful, idx_ok = 0.0, 0.0, True
    for li, i in enumerate(lanes):
        exact = orc_py.Pipeline(orc_py.config(64, voxel_stable=1), with_mapping=False)
        faithful = orc_py.Pipeline(orc_py.config(64, voxel_stable=0), with_mapping=False)
        for k in range(PARITY_STEPS):
            scan = pool[int(lane_ids(k, B, rank)[i])]
            pe, pf = exact.step(scan), faithful.step(scan)
            g = gpu[k][li]
            worst_exact = max(worst_exact, float(np.abs(g[4:7] - pe["t_odom"]).max()), float(np.abs(g[0:4] - pe["q_odom"]).max()))
            ang = 2 * np.arccos(min(1.0, abs(float(np.dot(g[0:4], pf["q_odom"])))))
            worst_faithful = max(worst_faithful, float(np.abs(g[4:7] - pf["t_odom"]).max()), float(ang))
        o = orc_py.extract_features(pool[int(lane_ids(PARITY_STEPS - 1, B, rank)[i])], orc_py.config(64, voxel_stable=1))
        for key in ("sharp_idx", "less_sharp_idx", "flat_idx"):
            idx_ok = idx_ok and np.array_equal(feats[i][key], o[key])
    ok = idx_ok and worst_exact < 1e-9 and worst_faithful < 1e-4
    return {"status": "ok" if ok else "FAILED", "lanes": lanes, "steps": PARITY_STEPS, "feature_indices_bit_exact": bool(idx_ok),
            "max_pose_err_vs_oracle_same_voxel_order": worst_exact, "max_pose_err_vs_reference_faithful_order": worst_faithful,
            "bars": {"same_order": 1e-9, "faithful": 1e-4}}


def run_ours(args, rank, world, local_rank):
    import torch
    ll = importlib.import_module("light-loam_b200")
    torch.cuda.set_device(local_rank)
    n_aff = pin_to_gpu_numa(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        quiet_nccl()
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    B = args.batch
    ctx = ll.Context(scan_line=64, batch=B, device=local_rank, max_ring_points=args.max_ring_points)
    pool = make_pool(ll, POOL_SCANS)
    ctx.pool_upload(pool)
    stream = torch.cuda.ExternalStream(ctx.cuda_stream(), device=local_rank)

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

    step = 0
    W = max(args.warmup, 3)
    # priming (state, not timing): the first frame only initialises (LO:427-431) and the graph vote starts at
    # now_frame > 5 (LO:794); the timed steps must run the steady-state path
    for _ in range(7 + W):
        ctx.process_pool(lane_ids(step, B, rank), want_poses=False)
        step += 1
    # ---- value: inputs resident in HBM (scan pool), device-timed ----------------------------------------------
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    launches = 0
    for _ in range(args.steps):
        ctx.process_pool(lane_ids(step, B, rank), want_poses=False)
        launches += ctx.L.ll_launch_count(ctx.h)
        step += 1
    e1.record(stream)
    torch.cuda.synchronize()
    sampler.stop_flag = True
    barrier()
    ms_total = maxreduce(e0.elapsed_time(e1))
    value = B * world * args.steps / (ms_total * 1e-3)
    st = ctx.stats()
    counts = {"n_raw": float(np.mean([len(p) for p in pool])), "n_full": st.n_full, "n_less_flat": st.n_less_flat,
              "n_less_sharp": st.n_less_sharp, "n_sharp": st.n_sharp, "n_flat": st.n_flat}

    # ---- e2e: host buffers through the public call, H2D of every scan + D2H of the poses in the timed region ---
    # One pinned host arena holds the scan sequence as packed xyz records (12 bytes per point: what scanRegistration
    # consumes, SR:105-110), with the first `B` scans repeated at the end so that the B scans of any step are one
    # contiguous block: ll_submit_packed moves a step's input with ONE host-to-device copy.
    n_pts = [len(p) for p in pool]
    seq = list(range(POOL_SCANS)) + list(range(B))
    offs = np.zeros(len(seq) + 1, np.int64)
    offs[1:] = np.cumsum([n_pts[j] * 12 for j in seq])
    arena_t = torch.empty(int(offs[-1]), dtype=torch.uint8).pin_memory()
    arena = arena_t.numpy()
    for q, j in enumerate(seq):
        arena[offs[q]:offs[q + 1]] = np.ascontiguousarray(pool[j][:, :3]).view(np.uint8).reshape(-1)
    cnts = np.array([n_pts[j] for j in seq], np.int32)
    views12 = [np.frombuffer(arena, dtype=np.float32, count=n_pts[j] * 3, offset=int(offs[q])).reshape(-1, 3) for q, j in enumerate(seq)]

    def first_slot(stp):
        return (lane_base(B, rank) + stp) % POOL_SCANS   # arena slots first_slot .. first_slot + B - 1 hold the step's scans

    for _ in range(3):
        f = first_slot(step)
        ctx.process_scans(views12[f:f + B])
        step += 1
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        f = first_slot(step)
        poses = ctx.process_scans(views12[f:f + B])
        step += 1
    torch.cuda.synchronize()
    e2e_sync_s = maxreduce(time.perf_counter() - t0)
    # the same through the asynchronous form of the call (ll_submit_packed / ll_collect, two submissions in flight):
    # the H2D copy of step k+1 overlaps the kernels of step k; every byte still crosses PCIe inside the timed region
    for k in range(W):   # warm-up of the asynchronous path (its second staging slab, copy stream and events are created on first use)
        f = first_slot(step)
        ctx.submit_packed(arena, offs[f:f + B], cnts[f:f + B], 12)
        step += 1
        if k > 0:
            ctx.collect()
    ctx.collect()
    barrier()
    h2d = 0
    t0 = time.perf_counter()
    for k in range(args.steps):
        f = first_slot(step)
        ctx.submit_packed(arena, offs[f:f + B], cnts[f:f + B], 12)
        h2d += int(offs[f + B] - offs[f])
        step += 1
        if k > 0:
            poses = ctx.collect()
    poses = ctx.collect()
    torch.cuda.synchronize()
    e2e_s = maxreduce(time.perf_counter() - t0)
    e2e_value = B * world * args.steps / e2e_s
    assert np.isfinite(poses).all()

    # ---- per-kernel device time (CUDA event pairs on the launching stream) -> roofline of the dominant kernel ----
    ctx.profile_enable(True)
    for _ in range(args.steps):
        ctx.process_pool(lane_ids(step, B, rank), want_poses=False)
        step += 1
    prof = ctx.profile_read()
    ctx.profile_enable(False)
    total_prof = sum(v[0] for v in prof.values())
    peak, peak_src = peaks()
    ab = algorithmic_bytes(counts)
    per_kernel = {}
    for name, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
        avg_ms = ms / n
        bytes_launch = ab.get(name, 0.0) * B
        per_kernel[name] = {"ms_per_launch": round(avg_ms, 4), "launches_per_step": n / args.steps, "share": round(ms / total_prof, 4),
                            "alg_gbs": round(bytes_launch / (avg_ms * 1e-3) / 1e9, 1) if bytes_launch else None}
    # the roofline object is that of the NN kernel (BASELINE.json: "NN + JtJ HBM GB/s vs peak"), which is also the kernel
    # with the largest share of the odometry stage
    dom = NN_KERNEL if NN_KERNEL in prof else max(prof.items(), key=lambda kv: kv[1][0])[0]
    dom_ms = prof[dom][0] / prof[dom][1]
    achieved = ab.get(dom, 0.0) * B / (dom_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        if tj.get("batch") == B and dom in tj.get("kernels", {}):
            traffic = tj["kernels"][dom]   # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture at this batch
            traffic_src = tj.get("source")
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 4),
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "alg_bytes_per_launch": int(ab.get(dom, 0.0) * B),
                "ms_per_launch": round(dom_ms, 4), "share_of_step": round(prof[dom][0] / total_prof, 4), "kernels": per_kernel}

    parity = None
    if rank == 0 and not args.no_parity:
        parity = parity_check(ll, ctx, pool, B, rank)

    # ---- CPU baseline: the oracle port of the same path on one host core, bounded sample (rank 0, N = 1) ---------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_sample(pool, n_scans=args.cpu_scans)

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "scans/s", "n_gpus": world, "steps": args.steps, "warmup": W,
            "ms_per_step": round(ms_total / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (features, NN) + f64 (residuals, LM)", "data": "synthetic",
            "config": config_dict(B, world, counts["n_raw"]), "host": {"cpu_affinity_cores": n_aff, "cores": os.cpu_count()},
            "e2e": {"value": round(e2e_value, 1), "unit": "scans/s", "h2d_bytes_per_step": int(h2d / args.steps), "d2h_bytes_per_step": B * (14 * 8 + 4),
                    "api": "ll_submit_packed / ll_collect (one pinned host arena of packed xyz records, one H2D copy per step, 2 submissions in flight)",
                    "h2d_gbs": round(h2d / e2e_s / 1e9, 2), "sync_call_value": round(B * world * args.steps / e2e_sync_s, 1)},
            "gpu_launches": int(launches), "roofline": roofline, "clocks": sampler.summary(), "parity": parity,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()
    if parity is not None and parity["status"] != "ok":
        sys.exit(3)


def cpu_baseline_sample(pool, n_scans):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc_py
    pipe = orc_py.Pipeline(orc_py.config(64), with_mapping=False)
    ms = []
    for k in range(n_scans + 2):
        r = pipe.step(pool[k % POOL_SCANS])
        if k >= 2:
            ms.append(r["ms"][:2].sum())
    per = float(np.mean(ms)) * 1e-3
    return {"value": round(1.0 / per, 2), "unit": "scans/s", "cores": 1, "kind": "port",
            "sample": "%d consecutive scans of the same pool, oracle extract_features + odometry (restated reference CPU path; PCL/Ceres unavailable offline), steady_clock per stage" % n_scans}


_G = {}


def _ref_worker(arg):
    wid, n_scans = arg
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc_py
    pipe = orc_py.Pipeline(orc_py.config(64), with_mapping=False)
    pool = _G["pool"]
    off = wid % POOL_SCANS
    for k in range(2):  # init frame + one warm frame, untimed
        pipe.step(pool[(off + k) % POOL_SCANS])
    t0 = time.perf_counter()
    for k in range(2, 2 + n_scans):
        pipe.step(pool[(off + k) % POOL_SCANS])
    return time.perf_counter() - t0


def run_reference(args, rank, world):
    """The reference's own CPU algorithm for the path (oracle restatement: the reference cannot be compiled here),
    one independent scan stream per host core, all cores."""
    if rank != 0:
        return
    import multiprocessing as mp
    ll = importlib.import_module("light-loam_b200")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import orc_py
    orc_py.build()
    cores = os.cpu_count() or 1
    _G["pool"] = make_pool(ll, POOL_SCANS)
    per_step = 2  # scans per worker per step: a bounded sample of the workload
    n_scans = per_step * (args.steps + max(args.warmup, 3))
    ctxm = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctxm.Pool(cores) as pl:
        times = pl.map(_ref_worker, [(w, n_scans) for w in range(cores)])
    wall = max(times)
    value = cores * n_scans / wall
    line = {"impl": "reference", "metric": METRIC, "value": round(value, 2), "unit": "scans/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(wall / (args.steps + max(args.warmup, 3)) * 1e3, 3), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 (features, NN) + f64 (residuals, LM)", "data": "synthetic",
            "config": config_dict(args.batch, world, float(np.mean([len(p) for p in _G["pool"]]))),
            "reference_note": "restated reference CPU path (PCL/Ceres/ROS unavailable offline), one scan stream per host core; the config object is the GPU arm's: same scans, same per-scan work",
            "cpu_baseline": {"value": round(value, 2), "unit": "scans/s", "cores": cores, "kind": "port",
                             "sample": "%d scans per core x %d cores, extract_features + odometry" % (n_scans, cores)},
            "e2e": {"value": round(value, 2), "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": round(time.perf_counter() - t0, 2)}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("LL_BENCH_BATCH", "256")), help="scan streams (lanes) per GPU")
    ap.add_argument("--cpu-scans", type=int, default=200, help="scans in the cpu_baseline sample")
    ap.add_argument("--config", type=int, default=2, choices=[1, 2, 3, 4, 5], help="BASELINE.json configuration (1-based): 2 = the headline (default); 1, 3, 4, 5 in bench_configs.py")
    ap.add_argument("--stream-scans", type=int, default=10000, help="--config 4: scans in the stream")
    ap.add_argument("--lanes", type=int, default=128, help="--config 4: sub-segments (lanes) per GPU")
    ap.add_argument("--overlap", type=int, default=6, help="--config 4: scans a sub-segment starts before its own first scan (1 = the anchor only)")
    ap.add_argument("--small", action="store_true", help="--config 5: 3e5-point map")
    ap.add_argument("--check", action="store_true", help="--config 5: compare with a single-GPU context")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the in-run oracle check of the bench configuration")
    ap.add_argument("--max-ring-points", type=int, default=6155, help="ring capacity (library default 6155; 3083 = 512-key sectors only)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.config != 2:
        import bench_configs
        {1: bench_configs.run_config1, 3: bench_configs.run_config3, 4: bench_configs.run_config4, 5: bench_configs.run_config5}[args.config](args, rank, world, local_rank)
    elif args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
