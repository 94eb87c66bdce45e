"""Host-side checks that need no GPU: the C-ABI library loads and exports every symbol include/*.h declares;
config defaults mirror the launch files; the multi-GPU sharding logic under gloo (world_size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(ll):
    hdr = open(os.path.join(ROOT, "include", "lightloam_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(ll_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 20
    L = ll.capi.lib()
    for sym in declared:
        assert hasattr(L, sym), sym
    assert declared == set(ll.capi.SYMBOLS)


def test_default_config_mirrors_launch_files(ll):
    c64, c16, c32 = (ll.default_config(n) for n in (64, 16, 32))
    assert (c64.minimum_range, round(c64.line_res, 3), round(c64.plane_res, 3)) == (5.0, 0.4, 0.8)   # aloam_velodyne_HDL_64.launch:8-12
    assert (round(c16.minimum_range, 3), round(c16.line_res, 3), round(c16.plane_res, 3)) == (0.3, 0.2, 0.4)
    assert round(c32.minimum_range, 3) == 0.3 and c64.graph_from_frame == 5
    assert abs(c64.lower_bound + 24.9) < 1e-6 and c64.up_bound == 2.0 and c64.batch == 1


def test_strerror_and_bad_arguments(ll):
    L = ll.capi.lib()
    assert L.ll_strerror(0) == b"ok" and b"capacity" in L.ll_strerror(-2)
    cfg = ll.default_config(64)
    cfg.scan_line = 48                         # SR:447-451: only 16 / 32 / 64
    h = ctypes.c_void_p()
    assert L.ll_create(ctypes.byref(cfg), ctypes.byref(h)) == ll.capi.LL_E_INVAL and not h.value
    assert L.ll_create(None, ctypes.byref(h)) == ll.capi.LL_E_INVAL


def test_synth_generator_is_deterministic_and_ring_consistent(ll):
    a, b = ll.synth.scan(16, 4), ll.synth.scan(16, 4)
    assert np.array_equal(a, b) and a.shape == (16000, 4)
    ang = np.degrees(np.arctan2(a[:, 2], np.hypot(a[:, 0], a[:, 1])))
    ring = np.round((ang + 15) / 2)
    assert np.abs(ang - (ring * 2 - 15)).max() < 1e-3 and set(ring.astype(int)) == set(range(16))


def test_two_rank_gloo_sharding(tmp_path):
    """world_size 2 on CPU (gloo): lanes of different ranks walk disjoint scan streams and the max-over-ranks
    reduction used for the timing is wired correctly."""
    code = r'''
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
import bench
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
ids = bench.lane_ids(5, 8, r)
allids = [torch.zeros(8, dtype=torch.int32) for _ in range(w)]
dist.all_gather(allids, torch.from_numpy(ids))
flat = torch.cat(allids).numpy()
assert len(set(flat.tolist())) == 16, flat
t = torch.tensor([float(r + 1)], dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.item() == w
nxt = bench.lane_ids(6, 8, r)
assert ((nxt - ids) %% bench.POOL_SCANS == 1).all()
dist.destroy_process_group()
print("ok", r)
''' % ROOT
    script = tmp_path / "gloo_shard_check.py"
    script.write_text(code)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29617", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.count("ok") == 2


def test_slab_bounds_partition_the_line(ll):
    mg = ll.multigpu
    for world in (1, 2, 3, 8):
        b = [mg.slab_bounds(r, world, -60.0, 60.0) for r in range(world)]
        assert b[0][0] == -np.inf and b[-1][1] == np.inf
        for r in range(world - 1):
            assert b[r][1] == b[r + 1][0]           # half-open slabs: every x has exactly one owner
    pts = np.random.default_rng(0).uniform(-60, 60, (1000, 4)).astype(np.float32)
    lo, hi = mg.slab_bounds(1, 4, -60.0, 60.0)
    sub = mg.slab_with_halo(pts, lo, hi, halo=1.0)
    assert sub[:, 0].min() >= lo - 1.0 and sub[:, 0].max() < hi + 1.0
    assert len(sub) == int(((pts[:, 0] >= lo - 1) & (pts[:, 0] < hi + 1)).sum())


def test_two_rank_gloo_handle_exchange(tmp_path):
    """The host-side plumbing of the multi-GPU scan-to-map mode under gloo, world_size 2: every rank ends up with all
    mailbox handles in rank order and with its own slab (a stub stands in for the CUDA context)."""
    code = r"""
import importlib, sys, torch.distributed as dist
sys.path.insert(0, %r)
mg = importlib.import_module("light-loam_b200.multigpu")
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
class Stub:
    def comm_export(self): return bytes([r]) * 64
    def comm_attach(self, rank, world, handles): self.got = (rank, world, handles)
    def map_set_slab(self, lo, hi): self.slab = (lo, hi)
s = Stub()
lo, hi = mg.attach_all(s, dist, -60.0, 60.0)
assert s.got[0] == r and s.got[1] == w and [h[0] for h in s.got[2]] == list(range(w)) and all(len(h) == 64 for h in s.got[2])
assert s.slab == (lo, hi) and ((r == 0 and hi == 0.0) or (r == 1 and lo == 0.0))
dist.destroy_process_group()
print("ok", r)
""" % ROOT
    script = tmp_path / "gloo_handles.py"
    script.write_text(code)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29618", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.count("ok") == 2


def test_two_rank_gloo_segment_chaining(tmp_path):
    """configs[3] exchange step under gloo, world_size 2: a stream of relative motions cut into two segments, each
    rank chaining its own increments from identity; after the 7-double all-gather + prefix product the global poses equal
    the single-process chain."""
    code = r"""
import importlib, sys, numpy as np, torch.distributed as dist
sys.path.insert(0, %r)
mg = importlib.import_module("light-loam_b200.multigpu")
dist.init_process_group("gloo")
r, w = dist.get_rank(), dist.get_world_size()
rng = np.random.default_rng(11)
n = 21
inc = []
for k in range(n):
    ax = rng.normal(0, 0.02, 3); ang = np.linalg.norm(ax)
    q = list(np.sin(ang / 2) * ax / ang) + [np.cos(ang / 2)]
    inc.append((q, list(rng.normal([1, 0, 0], 0.05))))
def chain(incs):
    q, t, out = [0, 0, 0, 1.0], [0, 0, 0.0], []
    for (dq, dt) in incs:                      # LO:830-831
        rt = mg.quat_rotate(q, dt); t = [t[i] + rt[i] for i in range(3)]; q = mg.quat_mul(q, dq)
        out.append(q + t)
    return np.array(out)
full = chain(inc)
b, e = mg.segment_ranges(n, w)[r]
assert mg.segment_ranges(n, w)[0][0] == 0 and mg.segment_ranges(n, w)[-1][1] == n
local = chain(inc[b:e])
glob = mg.chain_segments(local, dist)
assert np.abs(glob - full[b:e]).max() < 1e-12, np.abs(glob - full[b:e]).max()
dist.destroy_process_group()
print("ok", r)
""" % ROOT
    script = tmp_path / "gloo_chain.py"
    script.write_text(code)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29619", str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.count("ok") == 2


def test_config4_chain_prefix_and_anchor_algebra():
    """bench_configs (BASELINE.json configs[3]): the log-step prefix of the sub-segment transforms equals the sequential
    LO:830-831 chain, and inverse(anchor) o pose recovers the poses relative to a lane's anchor scan."""
    import bench_configs as bc
    rng = np.random.default_rng(4)
    K = 53
    q = rng.normal(size=(K, 4)) * 0.05
    q[:, 3] = 1
    q /= np.linalg.norm(q, axis=1)[:, None]
    ends = np.concatenate([q, rng.normal(size=(K, 3))], 1)
    out = bc.chain_prefix(ends)
    qq, tt = np.array([0, 0, 0, 1.0]), np.zeros(3)
    for k in range(K):
        assert np.allclose(out[k, :4], qq, atol=1e-13) and np.allclose(out[k, 4:], tt, atol=1e-12), k
        tt = tt + bc._qrot(qq, ends[k, 4:7])
        qq = bc._qmul(qq, ends[k, :4])
    # a lane's poses P[j] in its own frame; anchor a: rel[j] = inverse(P[a]) o P[j]; composing back gives P[j]
    P = np.concatenate([q[:10], rng.normal(size=(10, 3))], 1)
    rel = bc._qinv_apply(P[3][None, :], P)
    assert np.allclose(rel[3], [0, 0, 0, 1, 0, 0, 0], atol=1e-14)
    back_q = bc._qmul(P[3][None, :4], rel[:, :4])
    back_t = P[3][None, 4:] + bc._qrot(P[3][None, :4], rel[:, 4:])
    assert np.allclose(back_q, P[:, :4], atol=1e-13) and np.allclose(back_t, P[:, 4:], atol=1e-12)
