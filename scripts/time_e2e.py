"""End-to-end rate through ll_submit_scans / ll_collect with pinned host scans (the `e2e` leg of bench.py alone).
usage: [LL_B=256] [LL_STEPS=30] python scripts/time_e2e.py"""
import importlib, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ll = importlib.import_module("light-loam_b200")
B = int(os.environ.get("LL_B", "256")); steps = int(os.environ.get("LL_STEPS", "30")); NP = 157
ctx = ll.Context(scan_line=64, batch=B)
pool = [ll.synth.scan(64, k, mode=1) for k in range(NP)]
pinned = [torch.empty((len(p), 4), dtype=torch.float32).pin_memory() for p in pool]
for t, p in zip(pinned, pool):
    t.numpy()[:] = p
host = [t.numpy() for t in pinned]
views = ctx.make_views(host)
ids = lambda s: (((np.arange(B) * 7) + s) % NP).astype(np.int32)
for s in range(8):
    ctx.submit_views([views[i] for i in ids(s)])
    if s > 0:
        ctx.collect()
ctx.collect()
t0 = time.perf_counter()
for k in range(steps):
    ctx.submit_views([views[i] for i in ids(8 + k)])
    if k > 0:
        poses = ctx.collect()
poses = ctx.collect()
dt = time.perf_counter() - t0
nbytes = sum(host[i].nbytes for i in ids(8)) 
print("e2e scans/s %.0f  ms/step %.3f  H2D GB/s %.1f  pose0 %s" % (B * steps / dt, dt / steps * 1e3, nbytes * steps / dt / 1e9, np.array2string(poses[0][4:7], precision=5)))
