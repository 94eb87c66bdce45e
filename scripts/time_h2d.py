"""Where does the end-to-end rate go?  Raw pinned H2D bandwidth (one big copy vs many small ones, with and without
kernels running) next to the ll_submit_packed / ll_submit_scans rates.  usage: [LL_B=256] [LL_STEPS=20] python scripts/time_h2d.py"""
import importlib, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
ll = importlib.import_module("light-loam_b200")
B = int(os.environ.get("LL_B", "256")); steps = int(os.environ.get("LL_STEPS", "20"))
if os.environ.get("LL_PIN", "1") == "1":
    print("affinity cores", bench.pin_to_gpu_numa(0))
torch.cuda.set_device(0)
# --- raw copies ---------------------------------------------------------------------------------------------------
nbytes = 400 * 1000 * 1000
host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
s = torch.cuda.Stream()
def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
with torch.cuda.stream(s):
    t = timed(lambda: dev.copy_(host, non_blocking=True))
    print("raw 1 x 400 MB: %.1f GB/s" % (nbytes / t / 1e9))
    ch = nbytes // 256
    t = timed(lambda: [dev[i * ch:(i + 1) * ch].copy_(host[i * ch:(i + 1) * ch], non_blocking=True) for i in range(256)])
    print("raw 256 x 1.56 MB: %.1f GB/s" % (nbytes / t / 1e9))
    for mb in (4, 16, 64):
        c = mb * 1000 * 1000; n = nbytes // c
        t = timed(lambda: [dev[i * c:(i + 1) * c].copy_(host[i * c:(i + 1) * c], non_blocking=True) for i in range(n)])
        print("raw %d x %d MB: %.1f GB/s" % (n, mb, n * c / t / 1e9))
del host, dev
# --- through the library --------------------------------------------------------------------------------------------
ctx = ll.Context(scan_line=64, batch=B)
pool = bench.make_pool(ll, 64)
P = len(pool)
n_pts = [len(p) for p in pool]
seq = [j % P for j in range(P + B)]
offs = np.zeros(len(seq) + 1, np.int64); offs[1:] = np.cumsum([n_pts[j] * 12 for j in seq])
arena_t = torch.empty(int(offs[-1]), dtype=torch.uint8).pin_memory(); arena = arena_t.numpy()
for q, j in enumerate(seq):
    arena[offs[q]:offs[q + 1]] = np.ascontiguousarray(pool[j][:, :3]).view(np.uint8).reshape(-1)
cnts = np.array([n_pts[j] for j in seq], np.int32)
views12 = [np.frombuffer(arena, dtype=np.float32, count=n_pts[j] * 3, offset=int(offs[q])).reshape(-1, 3) for q, j in enumerate(seq)]
lv = ctx.make_views(views12)
def run(kind):
    st = 0
    def sub():
        nonlocal st
        f = st % P; st += 1
        if kind == "packed": ctx.submit_packed(arena, offs[f:f + B], cnts[f:f + B], 12)
        else: ctx.submit_views(lv[f:f + B])
    for k in range(4):
        sub()
        if k > 0: ctx.collect()
    ctx.collect()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for k in range(steps):
        sub()
        if k > 0: ctx.collect()
    ctx.collect(); torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print("%s: %.0f scans/s, %.2f ms/step, H2D %.1f GB/s" % (kind, B * steps / dt, dt / steps * 1e3, (offs[B] - offs[0]) * steps / dt / 1e9))
run("packed"); run("views"); run("packed")
# copy alone through the library's stream pattern: stage only
t0 = time.perf_counter()
for k in range(steps):
    ctx.stage_scans(views12[k % P:k % P + B]) if False else None
