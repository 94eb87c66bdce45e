"""Per-kernel device times of the fused pipeline (ll_profile_enable event pairs), B lanes from a resident scan pool.
usage: [LL_B=128] python scripts/prof_kernels.py [substring filter ...]"""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ll = importlib.import_module("light-loam_b200")
B = int(os.environ.get("LL_B", "128"))
ctx = ll.Context(scan_line=64, batch=B)
NP = 40
pool = [ll.synth.scan(64, k, mode=1) for k in range(NP)]
ctx.pool_upload(pool)
ids = lambda s: (((np.arange(B) * 7) + s) % NP).astype(np.int32)
for s in range(10):
    ctx.process_pool(ids(s), want_poses=False)
ctx.profile_enable(True)
for s in range(10, 18):
    ctx.process_pool(ids(s), want_poses=False)
prof = ctx.profile_read()
tot = sum(v[0] for v in prof.values()) / 8
flt = sys.argv[1:]
print("step %.3f ms |" % tot, " ".join("%s=%.4f" % (k.replace("k_", ""), v[0] / v[1]) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0]) if not flt or any(f in k for f in flt)))
