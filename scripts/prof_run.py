"""Short steady-state run for ncu: 64 lanes, 12 steps of the fused pipeline from a resident scan pool."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ll = importlib.import_module("light-loam_b200")
B = int(os.environ.get("LL_B", "64"))
steps = int(os.environ.get("LL_STEPS", "12"))
ctx = ll.Context(scan_line=64, batch=B)
pool = [ll.synth.scan(64, k, mode=1) for k in range(24)]
ctx.pool_upload(pool)
for s in range(steps):
    ids = ((np.arange(B) * 0 + s) % 24).astype(np.int32) if False else ((np.arange(B) % 8) + s) % 24
    ctx.process_pool(ids.astype(np.int32), want_poses=(s == steps - 1))
print("done", ctx.stats().kernel_launches)
