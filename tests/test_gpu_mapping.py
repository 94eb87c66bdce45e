"""GPU parity, scan-to-map (LM:1581-2168) vs the oracle: the fused pipeline with mapping on, and ll_mapping_step fed
with the oracle's own clouds.  Bar: mapped pose within 1e-4 m / 1e-4 rad per scan of the reference-faithful oracle
(north_star) and within 1e-7 of the oracle run in the GPU's voxel order."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _ang(qa, qb):
    return 2 * np.arccos(min(1.0, abs(float(np.dot(qa, qb)))))


@pytest.mark.parametrize("line,n,az", [(16, 10, None), (64, 8, None), (32, 7, 1200)])
def test_fused_pipeline_with_mapping_matches_oracle(ll, orc, line, n, az):
    ctx = ll.Context(scan_line=line, enable_mapping=1, map_capacity=1 << 18)
    exact = orc.Pipeline(orc.config(line, voxel_stable=1), with_mapping=True)
    faithful = orc.Pipeline(orc.config(line, voxel_stable=0), with_mapping=True)
    for k in range(n):
        scan = ll.synth.scan(line, k, az_steps=az)
        pg = ctx.process_scans([scan])[0]
        pe, pf = exact.step(scan), faithful.step(scan)
        assert np.abs(pg[4:7] - pe["t_odom"]).max() < 1e-9, k
        assert np.abs(pg[11:14] - pe["t_map"]).max() < 1e-7 and np.abs(pg[7:11] - pe["q_map"]).max() < 1e-7, (k, pg[7:], pe["q_map"], pe["t_map"])
        assert np.abs(pg[11:14] - pf["t_map"]).max() < 1e-4 and _ang(pg[7:11], pf["q_map"]) < 1e-4, k
    st = ctx.stats()
    assert st.map_surf > 50 and st.map_corner > 10 and st.map_surf_corr > 100
    ctx.close()


def test_mapping_step_api_matches_oracle(ll, orc):
    line = 16
    ocfg = orc.config(line, voxel_stable=1)
    ctx = ll.Context(scan_line=line, map_capacity=1 << 17)
    omap = orc.Mapping(ocfg)
    odo = orc.Odometry(ocfg)
    for k in range(8):
        f = orc.extract_features(ll.synth.scan(line, k), ocfg)
        po = odo.step(f["sharp"], f["less_sharp"], f["flat"], f["less_flat"])
        mo = omap.step(f["less_sharp"], f["less_flat"], po["q_w"], po["t_w"])
        mg = ctx.mapping_step(f["less_sharp"], f["less_flat"], po["q_w"], po["t_w"])
        assert np.abs(mg["t"] - mo["t"]).max() < 1e-7 and np.abs(mg["q"] - mo["q"]).max() < 1e-7, k
        st = ctx.stats()
        info = mo["info"]
        assert (st.map_corner, st.map_surf, st.stack_corner, st.stack_surf) == tuple(int(v) for v in info[1:5]), k
        assert mg["rc"] == (ll.capi.LL_W_FEW_CORRESPONDENCES if info[0] else 0)
        if not info[0]:
            assert (st.map_corner_corr, st.map_surf_corr) == (int(info[5]), int(info[6])), k
    ctx.close()


def test_map_preload_and_cube_shift(ll, orc):
    """A preloaded map (ll_map_insert) plus a pose far from the origin: the cube array must shift (LM:1596-1779) and
    the poses must keep matching the oracle."""
    line = 16
    ocfg = orc.config(line, voxel_stable=1)
    ctx = ll.Context(scan_line=line, map_capacity=1 << 17)
    omap = orc.Mapping(ocfg)
    rng = np.random.default_rng(1)
    corner = (rng.uniform(-60, 60, size=(3000, 4))).astype(np.float32)
    surf = (rng.uniform(-60, 60, size=(20000, 4))).astype(np.float32)
    corner[:, 2] = rng.uniform(-2, 10, 3000)
    surf[:, 2] = rng.uniform(-2, 10, 20000)
    ctx.map_insert(corner, surf)
    omap.insert(corner, surf)
    f = orc.extract_features(ll.synth.scan(line, 0), ocfg)
    for k, t in enumerate([[0, 0, 0], [130, -20, 3], [260, 40, 1], [-180, 10, 0]]):   # jumps of several cubes
        q = np.array([0, 0, np.sin(0.05 * k), np.cos(0.05 * k)])
        mo = omap.step(f["less_sharp"], f["less_flat"], q, np.array(t, float))
        mg = ctx.mapping_step(f["less_sharp"], f["less_flat"], q, np.array(t, float))
        assert np.abs(mg["t"] - mo["t"]).max() < 1e-7 and np.abs(mg["q"] - mo["q"]).max() < 1e-7, k
        st = ctx.stats()
        assert (st.map_corner, st.map_surf) == (int(mo["info"][1]), int(mo["info"][2])), (k, st.map_corner, st.map_surf, mo["info"])
    ctx.close()


def test_cpp_host_driver_matches_python_path(ll, tmp_path):
    """host/ll_run (C++ over the C ABI, KITTI-format trajectory file of LM:2284-2325) vs the same run through ctypes."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(ll.capi.LIB_PATH), "host", "ll_run")
    out = tmp_path / "traj.txt"
    subprocess.check_call([exe, "--lines", "16", "--scans", "6", "--mapping", "--out", str(out)])
    rows = np.loadtxt(out)
    assert rows.shape == (6, 12)
    ctx = ll.Context(scan_line=16, enable_mapping=1, map_capacity=1 << 20)
    H0 = None
    for k in range(6):
        p = ctx.process_scans([ll.synth.scan(16, k)])[0]
        x, y, z, w = p[7:11]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
        H = np.eye(4)
        H[:3, :3], H[:3, 3] = R, p[11:14]
        if H0 is None:
            H0 = H
        rel = (np.linalg.inv(H0) @ H)[:3, :].reshape(-1)
        assert np.abs(rel - rows[k]).max() < 5e-5          # fp32 matrices + 7 significant digits in the file
    ctx.close()


def test_config3_large_map_scan_to_map(ll, orc):
    """BASELINE.json configs[2] in small: HDL-64 scan against a large voxelised local map (here ~0.45 M points
    preloaded in the 5 x 5 x 3 cubes), graph vote off (the reference's mapping call site is commented out, LM:2057-2072)."""
    line = 64
    ocfg = orc.config(line, voxel_stable=1)
    rng = np.random.default_rng(3)
    n_s, n_c = 400000, 40000
    # ground + four walls of the synthetic room, jittered, then random vertical edges as corner map
    surf = np.zeros((n_s, 4), np.float32)
    u = rng.uniform(-1, 1, (n_s, 2))
    which = rng.integers(0, 5, n_s)
    surf[:, 0] = np.where(which == 1, 60, np.where(which == 2, -60, u[:, 0] * 60))
    surf[:, 1] = np.where(which == 3, 40, np.where(which == 4, -40, u[:, 1] * 40))
    surf[:, 2] = np.where(which == 0, -1.73, rng.uniform(-1.73, 13, n_s))
    surf[:, :3] += rng.normal(0, 0.01, (n_s, 3))
    corner = np.zeros((n_c, 4), np.float32)
    poles = rng.uniform(-55, 55, (200, 2))
    pid = rng.integers(0, 200, n_c)
    corner[:, 0], corner[:, 1] = poles[pid, 0], poles[pid, 1] * 0.7
    corner[:, 2] = rng.uniform(-1.7, 8, n_c)
    corner[:, :3] += rng.normal(0, 0.01, (n_c, 3))
    ctx = ll.Context(scan_line=line, map_capacity=1 << 20)
    omap = orc.Mapping(ocfg)
    ctx.map_insert(corner, surf)
    omap.insert(corner, surf)
    f = orc.extract_features(ll.synth.scan(line, 0, mode=1), ocfg)      # sensor at (25, 0), heading +y
    q0 = np.array([0, 0, np.sin(np.pi / 4), np.cos(np.pi / 4)])
    t0 = np.array([25.0, 0.0, 0.0])
    for k in range(2):
        mo = omap.step(f["less_sharp"], f["less_flat"], q0, t0)
        mg = ctx.mapping_step(f["less_sharp"], f["less_flat"], q0, t0)
        st = ctx.stats()
        assert (st.map_corner, st.map_surf) == (int(mo["info"][1]), int(mo["info"][2]))
        assert st.map_surf > 300000 or k > 0      # the first frame sees the whole preloaded map, then it is voxel-filtered
        assert np.abs(mg["t"] - mo["t"]).max() < 1e-6 and np.abs(mg["q"] - mo["q"]).max() < 1e-6, (k, mg, mo)
        assert (st.map_corner_corr, st.map_surf_corr) == (int(mo["info"][5]), int(mo["info"][6]))
    ctx.close()


def _config5_inputs(ll, line=32):
    """BASELINE.json configs[4] in small: an HDL-32 scan against a preloaded map."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("prof_map", os.path.join(os.path.dirname(os.path.dirname(__file__)), "scripts", "prof_map.py"))
    pm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pm)
    corner, surf = pm.config3_map(n_s=300000, n_c=30000, seed=7)
    return corner, surf


def test_split_solve_matches_single_cta(ll, monkeypatch):
    """The LM solve split over 16 / 5 CTAs (in-kernel all-reduce through the mailbox) and over a thread-block cluster of 8
    (sums through distributed shared memory: the default on one GPU) vs one CTA: same pose to fp64 summation-order noise,
    same iteration counts and termination."""
    corner, surf = _config5_inputs(ll)
    q0 = np.array([0, 0, np.sin(np.pi / 4), np.cos(np.pi / 4)])
    t0 = np.array([25.0, 0.0, 0.0])
    res = {}
    for parts in (1, 16, 5, "cluster8", "cluster4"):
        if isinstance(parts, int):
            monkeypatch.setenv("LL_LM_PARTS", str(parts))
        else:
            monkeypatch.delenv("LL_LM_PARTS", raising=False)
            monkeypatch.setenv("LL_LM_CLUSTER", parts[7:])
        ctx = ll.Context(scan_line=32, map_capacity=1 << 19)
        f = ctx.extract_features(ll.synth.scan(32, 0, mode=1))
        ctx.map_insert(corner, surf)
        out = []
        for k in range(2):
            m = ctx.mapping_step(f["less_sharp"], f["less_flat"], q0, t0)
            st = ctx.stats()
            out.append((m["q"].copy(), m["t"].copy(), list(st.map_jacobian_evals), list(st.map_termination), st.map_corner_corr, st.map_surf_corr))
        res[parts] = out
        ctx.close()
    for parts in (16, 5, "cluster8", "cluster4"):
        for a, b in zip(res[1], res[parts]):
            assert np.abs(a[0] - b[0]).max() < 1e-11 and np.abs(a[1] - b[1]).max() < 1e-10, (parts, a, b)
            assert a[2:] == b[2:], (parts, a[2:], b[2:])
    assert res[1][0][5] > 1000


def test_two_contexts_slab_sharded_allreduce(ll):
    """Config 5 on one GPU: two contexts (= two ranks) with the same map, the stack points shared out by x slab, the 28
    doubles all-reduced inside the LM kernels through each other's mailboxes.  Both ranks must return bit-identical
    poses, equal to the single-context result to fp64 summation-order noise; every correspondence has one owner."""
    import threading
    corner, surf = _config5_inputs(ll)
    q0 = np.array([0, 0, np.sin(np.pi / 4), np.cos(np.pi / 4)])
    t0 = np.array([25.0, 0.0, 0.0])
    single = ll.Context(scan_line=32, map_capacity=1 << 19)
    f = single.extract_features(ll.synth.scan(32, 0, mode=1))
    single.map_insert(corner, surf)
    want = single.mapping_step(f["less_sharp"], f["less_flat"], q0, t0)
    st1 = single.stats()
    ranks = [ll.Context(scan_line=32, map_capacity=1 << 19) for _ in range(2)]
    ptrs = [c.comm_local_ptr() for c in ranks]
    for r, c in enumerate(ranks):
        c.comm_attach(r, 2, ptrs)
        c.map_set_slab(*ll.multigpu.slab_bounds(r, 2, 10.0, 40.0))     # the sensor sits at x = 25: both slabs get work
        c.map_insert(corner, surf)
    got = [None, None]

    def run(r):
        got[r] = ranks[r].mapping_step(f["less_sharp"], f["less_flat"], q0, t0)

    th = [threading.Thread(target=run, args=(r,)) for r in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join(timeout=120)
    assert got[0] is not None and got[1] is not None
    assert got[0]["rc"] == 0 and got[1]["rc"] == 0
    assert np.array_equal(got[0]["q"], got[1]["q"]) and np.array_equal(got[0]["t"], got[1]["t"])
    assert np.abs(got[0]["t"] - want["t"]).max() < 1e-10 and np.abs(got[0]["q"] - want["q"]).max() < 1e-11
    s = [c.stats() for c in ranks]
    assert s[0].map_surf_corr > 100 and s[1].map_surf_corr > 100
    assert s[0].map_surf_corr + s[1].map_surf_corr == st1.map_surf_corr
    assert s[0].map_corner_corr + s[1].map_corner_corr == st1.map_corner_corr
    assert list(s[0].map_jacobian_evals) == list(st1.map_jacobian_evals) == list(s[1].map_jacobian_evals)
    for c in ranks + [single]:
        c.close()


def test_multi_process_config5_when_two_gpus(ll, tmp_path):
    """The same through CUDA IPC between processes, one GPU each (skipped on a 1-GPU box; run with gpurun --gpus 2)."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                          "--master-port", "29655", os.path.join(root, "scripts", "bench_config5.py"), "--check", "--small"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-3000:]
    assert "config5 ok" in out.stdout


def test_scan_to_map_graph_vote_matches_oracle(ll, orc):
    """BASELINE.json configs[2] "graph matching on": the vote the reference keeps commented out at LM:2057-2072 (LM's own
    copy of vote_simple, LM:836-1027: 20 regions, threshold 0.95, votes < 0.75 m selected, selected blocks added twice),
    switched on from the first mapping frame (map_graph_vote = 1).  Selected counts equal the oracle's, poses agree."""
    line = 16
    ocfg = orc.config(line, voxel_stable=1, map_graph_vote=1)
    ctx = ll.Context(scan_line=line, map_capacity=1 << 17, map_graph_vote=1)
    plain = ll.Context(scan_line=line, map_capacity=1 << 17)
    omap = orc.Mapping(ocfg)
    odo = orc.Odometry(ocfg)
    differs = False
    for k in range(8):
        f = orc.extract_features(ll.synth.scan(line, k), ocfg)
        po = odo.step(f["sharp"], f["less_sharp"], f["flat"], f["less_flat"])
        mo = omap.step(f["less_sharp"], f["less_flat"], po["q_w"], po["t_w"])
        mg = ctx.mapping_step(f["less_sharp"], f["less_flat"], po["q_w"], po["t_w"])
        mp = plain.mapping_step(f["less_sharp"], f["less_flat"], po["q_w"], po["t_w"])
        assert np.abs(mg["t"] - mo["t"]).max() < 1e-7 and np.abs(mg["q"] - mo["q"]).max() < 1e-7, k
        st, info = ctx.stats(), mo["info"]
        if not info[0]:
            assert st.map_surf_corr == int(info[6]) and st.map_vote_corr == int(info[6]), k
            assert st.map_vote_selected == int(info[7]), (k, st.map_vote_selected, int(info[7]))
            assert 0 < st.map_vote_selected <= st.map_vote_corr
            differs |= bool(np.abs(mg["t"] - mp["t"]).max() > 1e-12)
    assert differs          # the doubled blocks do change the solve
    ctx.close()
    plain.close()


def test_cpp_host_driver_reads_kitti_bin_files(ll, tmp_path):
    """ll_run --bin-dir: KITTI-style .bin files (float32 x, y, z, i records - what kittiHelper.cpp:22-32, 137-147 reads)
    give the same trajectory file as the generator path fed the same scans."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.abspath(ll.__file__)), "host", "ll_run")
    bins = tmp_path / "velodyne"
    bins.mkdir()
    for k in range(5):
        s = ll.synth.scan(16, k).copy()
        s[:, 3] = 0.25                                   # reflectance column: ignored by scanRegistration
        s.astype(np.float32).tofile(bins / ("%06d.bin" % k))
    a, b = tmp_path / "a.txt", tmp_path / "b.txt"
    subprocess.check_call([exe, "--lines", "16", "--scans", "5", "--out", str(a)])
    subprocess.check_call([exe, "--lines", "16", "--scans", "9", "--bin-dir", str(bins), "--out", str(b)])   # stops at the first missing file
    ra, rb = np.loadtxt(a), np.loadtxt(b)
    assert ra.shape == rb.shape == (5, 12)
    assert np.array_equal(ra, rb)


@pytest.mark.parametrize("mapping", [0, 1])
def test_graph_replay_equals_eager_launches(ll, monkeypatch, mapping):
    """After two eager calls a context with few lanes replays the step from a CUDA graph (with mapping: one graph per current
    cube-map buffer).  Same kernels, same arguments: the poses must be identical to the bit with LL_GRAPH=0."""
    line, n = 16, 9
    scans = [ll.synth.scan(line, k) for k in range(n)]
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("LL_GRAPH", mode)
        ctx = ll.Context(scan_line=line, batch=2, enable_mapping=mapping, map_capacity=1 << 18)
        out[mode] = np.stack([ctx.process_scans([scans[k], scans[(k + 3) % n]]) for k in range(n)])
        ctx.close()
    assert np.array_equal(out["1"], out["0"])
    assert np.abs(out["1"][-1, 0, 4:7]).max() > 0.1     # the trajectory moved


@pytest.mark.parametrize("n,key_bits,cap_extra", [(0, 32, 10), (1, 64, 0), (2047, 31, 1), (2048, 44, 0), (2049, 48, 4097), (100003, 56, 50000), (700001, 64, 0)])
def test_radix_sort_and_prefix_sum_primitives(ll, n, key_bits, cap_extra):
    """csrc/ll_sort.cuh on its own (the hand-written replacements of cub::DeviceRadixSort / DeviceScan in the map filter): random
    keys with many duplicates, sizes around the 2048-element tiles, arrays larger than the data (device-side lengths).
    Bar: bit-exact against numpy's stable sort / cumsum."""
    rng = np.random.default_rng(n + key_bits)
    ctx = ll.Context(scan_line=16)
    mask = np.uint64((1 << key_bits) - 1) if key_bits < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
    keys = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    keys &= mask
    if n > 10:
        keys[rng.integers(0, n, size=n // 2)] = keys[rng.integers(0, n, size=n // 2)]     # duplicates: stability matters
        keys[: n // 8] &= np.uint64(0xFF)                                                # and a crowd in the low digit only
    vals = np.arange(n, dtype=np.int32)
    scan_in = rng.integers(0, 3, size=n + 5, dtype=np.int32)
    k, v, sc = ctx.debug_sort_scan(keys, vals, key_bits=key_bits, capacity=n + 5 + cap_extra, scan=scan_in)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order]) and np.array_equal(v, vals[order])
    assert np.array_equal(sc, np.concatenate([[0], np.cumsum(scan_in[:-1])]).astype(np.int32))
    ctx.close()


def test_long_trajectory_with_mapping_stays_on_the_oracle(ll, orc):
    """30 scans with scan-to-map on (the map grows, cubes fill, the graph replays alternate between the two cube-map buffers):
    mapped pose within 1e-6 of the oracle in the same voxel order at every scan, map sizes equal at the end."""
    line, n = 16, 30
    ctx = ll.Context(scan_line=line, enable_mapping=1, map_capacity=1 << 18)
    exact = orc.Pipeline(orc.config(line, voxel_stable=1), with_mapping=True)
    for k in range(n):
        scan = ll.synth.scan(line, k, mode=1)
        pg = ctx.process_scans([scan])[0]
        pe = exact.step(scan)
        assert np.abs(pg[4:7] - pe["t_odom"]).max() < 1e-8, k
        assert np.abs(pg[11:14] - pe["t_map"]).max() < 1e-6 and np.abs(pg[7:11] - pe["q_map"]).max() < 1e-6, (k, pg[7:], pe["q_map"], pe["t_map"])
    st = ctx.stats()
    assert st.frame == n and st.map_surf > 1000
    ctx.close()
