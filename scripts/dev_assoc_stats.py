"""Development statistics of the NN thread pass: candidates per query by phase and range class.
Needs the instrumented build:  nvcc ... -DLL_ASSOC_STATS -o build_stats/liblightloam_stats.so  (profiling notes)."""
import ctypes, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
ll = importlib.import_module("light-loam_b200")
ll.capi.LIB_PATH = os.path.join(ROOT, "build_stats", "liblightloam_stats.so")
import bench
B = 64
ctx = ll.Context(scan_line=64, batch=B)
pool = bench.make_pool(ll, 96)
ctx.pool_upload(pool)
L = ll.capi.lib()
out = (ctypes.c_ulonglong * 32)()
for s in range(8):
    ctx.process_pool(((np.arange(B) + s) % 96).astype(np.int32))
L.ll_dev_assoc_stats(out, 1)
for s in range(8, 12):
    ctx.process_pool(((np.arange(B) + s) % 96).astype(np.int32))
L.ll_dev_assoc_stats(out, 0)
v = np.array(list(out), dtype=np.float64)
for kind, base in (("corner", 0), ("plane", 16)):
    print(kind, "heavy-queued", int(v[base + 15]))
    for name, o in (("rho<10", 0), ("10<=rho<20", 5), ("rho>=20", 10)):
        q = v[base + o]
        if q == 0:
            continue
        print("   %-11s queries %8d  per query: NN seed %6.1f  NN loop %6.1f  window first %6.1f  window loop %6.1f  total %6.1f" %
              (name, q, v[base + o + 1] / q, v[base + o + 2] / q, v[base + o + 3] / q, v[base + o + 4] / q, v[base + o + 1:base + o + 5].sum() / q))
