"""Wall-clock time per fused-pipeline step with the scan pool resident (no per-kernel profiling): the figure bench.py
reports as `value`, without its other legs.  usage: [LL_B=256] [LL_STEPS=40] python scripts/time_steps.py"""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ll = importlib.import_module("light-loam_b200")
B = int(os.environ.get("LL_B", "256")); steps = int(os.environ.get("LL_STEPS", "40")); NP = 157
ctx = ll.Context(scan_line=64, batch=B)
ctx.pool_upload([ll.synth.scan(64, k, mode=1) for k in range(NP)])
ids = lambda s: (((np.arange(B) * 7) + s) % NP).astype(np.int32)
for s in range(12):
    ctx.process_pool(ids(s), want_poses=(s == 11))
t0 = time.perf_counter()
for s in range(12, 12 + steps):
    poses = ctx.process_pool(ids(s), want_poses=(s == 11 + steps))
dt = time.perf_counter() - t0
print("ms/step %.4f  scans/s %.0f  pose0 %s" % (dt / steps * 1e3, B * steps / dt, np.array2string(poses[0][4:7], precision=6)))
