"""Experiment: the 256 lanes of the bench as C contexts of 256 / C lanes, each on its own stream, driven from one thread:
kernels of different contexts overlap (features of one with the odometry of another).  usage: python scripts/exp_two_ctx.py"""
import importlib, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
ll = importlib.import_module("light-loam_b200")
B, steps = 256, int(os.environ.get("LL_STEPS", "30"))
pool = bench.make_pool(ll, bench.POOL_SCANS)
for C in (1, 2, 4):
    ctxs = [ll.Context(scan_line=64, batch=B // C) for _ in range(C)]
    for c in ctxs:
        c.pool_upload(pool)
    ids = lambda s: bench.lane_ids(s, B, 0)
    def step_all(s):
        i = ids(s)
        for k, c in enumerate(ctxs):
            c.process_pool(i[k * (B // C):(k + 1) * (B // C)], want_poses=False)
    for s in range(10):
        step_all(s)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(10, 10 + steps):
        step_all(s)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("contexts %d x %d lanes: %.0f scans/s, %.3f ms per 256 scans" % (C, B // C, B * steps / dt, dt / steps * 1e3))
    for c in ctxs:
        c.close()
