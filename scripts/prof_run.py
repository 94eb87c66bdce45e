"""Short steady-state run for ncu: LL_B lanes (default 256), LL_STEPS steps (default 10) of the fused pipeline from the bench's
own resident scan pool (314 distinct scans, every lane reads its own: DRAM traffic under ncu is that of the bench)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
ll = importlib.import_module("light-loam_b200")
B = int(os.environ.get("LL_B", "256"))
steps = int(os.environ.get("LL_STEPS", "10"))
ctx = ll.Context(scan_line=64, batch=B)
ctx.pool_upload(bench.make_pool(ll, bench.POOL_SCANS))
for s in range(steps):
    ctx.process_pool(bench.lane_ids(s, B, 0), want_poses=(s == steps - 1))
print("done", ctx.stats().kernel_launches)
