"""GPU parity, scan-to-scan odometry (LO:425-896 incl. graph vote) vs the oracle on identical clouds.
Bar: association indices bit-exact; per-scan pose within 1e-4 m / 1e-4 rad of the reference-faithful oracle
(north_star), and within 1e-9 of the oracle run in the GPU's voxel order (same inputs to the last bit)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _ang(qa, qb):
    d = abs(float(np.dot(qa, qb)))
    return 2 * np.arccos(min(1.0, d))


@pytest.mark.parametrize("line,n,az,lm_threads", [(16, 10, None, None), (64, 9, None, None), (32, 8, 1200, None), (64, 8, None, "256"),
                                                  (64, 9, None, "many-lanes"), (16, 8, None, "many-lanes"), (64, 8, None, "cluster2")])
def test_fused_pipeline_trajectory_matches_oracle(ll, orc, line, n, az, lm_threads, monkeypatch):
    # one lane runs the single-stream forms by default: a warp per query (k_odom_assoc_direct) and the solve spread over a
    # thread-block cluster of 8, records + vote in one kernel; "many-lanes" forces the forms the batched path uses (thread pass +
    # queue, separate record and vote kernels, one CTA per solve)
    if lm_threads == "many-lanes":
        monkeypatch.setenv("LL_ASSOC_DIRECT", "0")
        monkeypatch.setenv("LL_LM_CLUSTER", "1")
        monkeypatch.setenv("LL_VOTE_FUSED", "0")
    elif lm_threads == "cluster2":
        monkeypatch.setenv("LL_LM_CLUSTER", "2")
    elif lm_threads:   # the CTA shape the solve uses when there are more scan streams than SMs (bench: 256 lanes)
        monkeypatch.setenv("LL_LM_THREADS", lm_threads)
        monkeypatch.setenv("LL_LM_CLUSTER", "1")
    ctx = ll.Context(scan_line=line)
    exact = orc.Pipeline(orc.config(line, voxel_stable=1), with_mapping=False)
    faithful = orc.Pipeline(orc.config(line, voxel_stable=0), with_mapping=False)
    for k in range(n):
        scan = ll.synth.scan(line, k, az_steps=az)
        pg = ctx.process_scans([scan])[0]
        pe, pf = exact.step(scan), faithful.step(scan)
        assert np.abs(pg[4:7] - pe["t_odom"]).max() < 1e-9 and np.abs(pg[0:4] - pe["q_odom"]).max() < 1e-9, k
        assert np.abs(pg[4:7] - pf["t_odom"]).max() < 1e-4 and _ang(pg[0:4], pf["q_odom"]) < 1e-4, k
    st = ctx.stats()
    assert st.frame == n and st.plane_selected[2] <= st.plane_corr[2] and st.corner_corr[2] > 50
    if n > 7:
        assert st.plane_selected[2] < st.plane_corr[2]      # the vote removed something once now_frame > 5
    ctx.close()


def test_committed_golden_trajectory(ll):
    g = np.load(os.path.join(GOLD, "trajectory_vlp16_10.npz"))["poses"]
    ctx = ll.Context(scan_line=16)
    for k in range(10):
        pg = ctx.process_scans([ll.synth.scan(16, k)])[0]
        assert np.abs(pg[0:7] - g[k][0:7]).max() < 1e-9, k
    ctx.close()


def test_association_indices_and_host_cloud_api(ll, orc):
    """ll_odometry_step fed with the oracle's own feature clouds (what the ROS node receives from the topics):
    correspondences (closest, 2nd, 3rd point) must be identical index for index, then the pose."""
    line = 64
    ctx = ll.Context(scan_line=line)
    ocfg = orc.config(line, voxel_stable=1)
    odo = orc.Odometry(ocfg)
    for k in range(8):
        f = orc.extract_features(ll.synth.scan(line, k), ocfg)
        po = odo.step(f["sharp"], f["less_sharp"], f["flat"], f["less_flat"])
        pg = ctx.odometry_step(f["sharp"], f["less_sharp"], f["flat"], f["less_flat"])
        assert np.abs(pg["t_w"] - po["t_w"]).max() < 1e-9 and np.abs(pg["q_w"] - po["q_w"]).max() < 1e-9
        assert np.abs(pg["t_last"] - po["t_last"]).max() < 1e-9
        if k == 0:
            continue
        oc, op = odo.assoc(len(f["sharp"]), len(f["flat"]))
        gc, gp = ctx.debug_assoc(0)
        gc, gp = gc[:len(f["sharp"])], gp[:len(f["flat"])]
        got_c = np.array([[i, a, b] for i, (a, b) in enumerate(gc) if b >= 0], np.int32).reshape(-1, 3)
        got_p = np.array([[i, a, b, c] for i, (a, b, c, _) in enumerate(gp) if a >= 0], np.int32).reshape(-1, 4)
        assert np.array_equal(got_c, oc), k
        assert np.array_equal(got_p, op), k
        stats = odo.stats()
        st = ctx.stats()
        assert [int(s[0]) for s in stats] == list(st.corner_corr) and [int(s[1]) for s in stats] == list(st.plane_corr)
        assert [int(s[2]) for s in stats] == list(st.plane_selected)
        assert [int(s[5]) for s in stats] == list(st.lm_jacobian_evals) and [int(s[7]) for s in stats] == list(st.lm_termination)
        assert np.allclose([s[4] for s in stats], list(st.final_cost), rtol=1e-9, atol=1e-12)
    ctx.close()


def test_lanes_are_independent_streams(ll):
    """batch = 4: lane i walking its own scan sequence equals a single-lane context fed the same sequence."""
    line, B, n = 16, 4, 6
    multi = ll.Context(scan_line=line, batch=B)
    singles = [ll.Context(scan_line=line) for _ in range(B)]
    for k in range(n):
        scans = [ll.synth.scan(line, k + 3 * i) for i in range(B)]
        pm = multi.process_scans(scans)
        for i in range(B):
            ps = singles[i].process_scans([scans[i]])[0]
            assert np.array_equal(pm[i], ps), (k, i)
    multi.close()
    for s in singles:
        s.close()


def test_pool_path_equals_host_path_and_reset(ll):
    line = 16
    scans = [ll.synth.scan(line, k) for k in range(5)]
    a = ll.Context(scan_line=line, batch=2)
    b = ll.Context(scan_line=line, batch=2)
    b.pool_upload(scans)
    for k in range(4):
        pa = a.process_scans([scans[k], scans[k + 1]])
        pb = b.process_pool([k, k + 1])
        assert np.array_equal(pa, pb)
    b.reset()
    first = b.process_pool([0, 1])
    assert np.array_equal(first[:, 4:7], np.zeros((2, 3)))      # state forgotten: first frame only initialises
    a.close()
    b.close()


def test_too_few_correspondences_warning(ll):
    """A second scan with nothing near the first one: no correspondences -> pose unchanged, warning code."""
    ctx = ll.Context(scan_line=16)
    f = ctx.extract_features(ll.synth.scan(16, 0))
    r0 = ctx.odometry_step(f["sharp"], f["less_sharp"], f["flat"], f["less_flat"])
    far = {k: v.copy() for k, v in f.items()}
    for k in ("sharp", "less_sharp", "flat", "less_flat"):
        far[k][:, :3] += 500.0
    r1 = ctx.odometry_step(far["sharp"], far["less_sharp"], far["flat"], far["less_flat"])
    assert r1["rc"] == ll.capi.LL_W_FEW_CORRESPONDENCES
    assert np.array_equal(r1["t_w"], r0["t_w"]) and np.array_equal(r1["q_last"], [0, 0, 0, 1])
    ctx.close()


def test_non_monotone_last_cloud_takes_literal_path(ll, orc):
    """Clouds that are NOT ring-sorted (never produced by scanRegistration, but legal input on the topic) must go
    through the literal scan loops of LO:504-553 / LO:668-721 and still match the oracle index for index."""
    line = 16
    ctx = ll.Context(scan_line=line)
    ocfg = orc.config(line, voxel_stable=1)
    odo = orc.Odometry(ocfg)
    rng = np.random.default_rng(5)
    for k in range(4):
        f = orc.extract_features(ll.synth.scan(line, k), ocfg)
        ls = f["less_sharp"][rng.permutation(len(f["less_sharp"]))]
        lf = f["less_flat"].copy()
        blocks = np.array_split(np.arange(len(lf)), 7)          # shuffle coarse blocks: locally sorted, globally not
        lf = lf[np.concatenate([blocks[i] for i in rng.permutation(7)])]
        po = odo.step(f["sharp"], ls, f["flat"], lf)
        pg = ctx.odometry_step(f["sharp"], ls, f["flat"], lf)
        assert np.abs(pg["t_w"] - po["t_w"]).max() < 1e-9 and np.abs(pg["q_w"] - po["q_w"]).max() < 1e-9, k
        if k:
            oc, op = odo.assoc(len(f["sharp"]), len(f["flat"]))
            gc, gp = ctx.debug_assoc(0)
            got_c = np.array([[i, a, b] for i, (a, b) in enumerate(gc[:len(f["sharp"])]) if b >= 0], np.int32).reshape(-1, 3)
            got_p = np.array([[i, a, b, c] for i, (a, b, c, _) in enumerate(gp[:len(f["flat"])]) if a >= 0], np.int32).reshape(-1, 4)
            assert np.array_equal(got_c, oc) and np.array_equal(got_p, op), k
    ctx.close()


def test_async_submit_collect_equals_sync_call(ll):
    line, B = 16, 3
    a = ll.Context(scan_line=line, batch=B)
    b = ll.Context(scan_line=line, batch=B)
    seqs = [[ll.synth.scan(line, k + 2 * i) for i in range(B)] for k in range(6)]
    want = [a.process_scans(s) for s in seqs]
    got = []
    b.submit_scans(seqs[0])
    for k in range(1, 6):
        b.submit_scans(seqs[k])
        got.append(b.collect())
    got.append(b.collect())
    for w, g in zip(want, got):
        assert np.array_equal(w, g)
    with pytest.raises(ll.LightLoamError):
        b.collect()                      # nothing outstanding
    a.close()
    b.close()


def _scene_cloud(rng, n, ring_map):
    """Points sampled in CARTESIAN space (ground, walls, a cluster hugging the sensor axis, a wedge across the +-pi
    azimuth seam): nothing like the even (azimuth, ring) spread of a spinning sensor.  Ring ids come from the
    elevation (64-line formula, clipped: rings 0 and 63 get wide bands), optionally scrambled by ring_map."""
    parts = []
    g = np.c_[rng.uniform(-40, 40, n // 2), rng.uniform(-40, 40, n // 2), rng.normal(-1.7, 0.02, n // 2)]
    parts.append(g)
    w = np.c_[np.full(n // 6, 18.0) + rng.normal(0, 0.02, n // 6), rng.uniform(-30, 30, n // 6), rng.uniform(-1.7, 6, n // 6)]
    parts.append(w)
    axis = np.c_[rng.normal(0, 0.3, n // 8), rng.normal(0, 0.3, n // 8), rng.uniform(2, 8, n // 8)]
    parts.append(axis)
    seam = np.c_[rng.uniform(-25, -6, n // 6), rng.normal(0, 0.4, n // 6), rng.uniform(-1.7, 2, n // 6)]
    parts.append(seam)
    p = np.concatenate(parts).astype(np.float32)
    el = np.degrees(np.arctan2(p[:, 2], np.hypot(p[:, 0], p[:, 1])))
    ring = np.clip(np.floor((el + 24.9) * (63.0 / 26.9) + 0.5), 0, 63).astype(np.int64)
    ring = ring_map[ring]
    order = np.argsort(ring, kind="stable")
    p, ring = p[order], ring[order]
    return np.c_[p, ring + rng.uniform(0, 0.09, len(p))].astype(np.float32)


@pytest.mark.parametrize("case", ["ordered_bands", "scrambled_rings"])
def test_polar_search_exact_on_non_sensor_like_clouds(ll, orc, case):
    """The exact 1-NN / ring-window search prunes with the rings' elevation bands and azimuth offsets.  Clouds whose
    geometry has nothing to do with a spinning sensor (Cartesian sampling, points on the sensor axis, a cluster
    across the azimuth seam, ring ids that do not follow the elevation, queries with nothing within 5 m) must still
    give the oracle's kd-tree + literal-loop correspondences index for index."""
    line = 64
    rng = np.random.default_rng(11 if case == "ordered_bands" else 12)
    ring_map = np.arange(64) if case == "ordered_bands" else (np.arange(64) * 27) % 64   # a permutation of the rings
    ctx = ll.Context(scan_line=line)
    ocfg = orc.config(line, voxel_stable=1)
    odo = orc.Odometry(ocfg)
    ls = _scene_cloud(rng, 6000, ring_map)
    lf = _scene_cloud(rng, 40000, ring_map)

    def queries(tgt, m):
        near = tgt[rng.choice(len(tgt), m - m // 8, replace=False)].copy()
        near[:, :3] += rng.normal(0, 0.08, (len(near), 3)).astype(np.float32)
        far = np.c_[rng.uniform(-60, 60, m // 8), rng.uniform(-60, 60, m // 8), rng.uniform(8, 30, m // 8), np.zeros(m // 8)].astype(np.float32)
        return np.concatenate([near, far]).astype(np.float32)

    for k in range(3):
        sharp, flat = queries(ls, 700), queries(lf, 1500)
        po = odo.step(sharp, ls, flat, lf)
        pg = ctx.odometry_step(sharp, ls, flat, lf)
        if k == 0:
            continue
        oc, op = odo.assoc(len(sharp), len(flat))
        gc, gp = ctx.debug_assoc(0)
        got_c = np.array([[i, a, b] for i, (a, b) in enumerate(gc[:len(sharp)]) if b >= 0], np.int32).reshape(-1, 3)
        got_p = np.array([[i, a, b, c] for i, (a, b, c, _) in enumerate(gp[:len(flat)]) if a >= 0], np.int32).reshape(-1, 4)
        assert len(oc) > 50 and len(op) > 100, (len(oc), len(op))
        assert np.array_equal(got_c, oc), k
        assert np.array_equal(got_p, op), k
        assert np.abs(pg["t_w"] - po["t_w"]).max() < 1e-7 and np.abs(pg["q_w"] - po["q_w"]).max() < 1e-7, k
    ctx.close()


def test_odometry_step_reads_pointxyzi_payloads_in_place(ll, orc):
    """Feature clouds as they travel between the nodes (PointCloud2, pcl::PointXYZI: point_step 32, intensity at byte 16)
    must give the same poses as packed float4 clouds."""
    line = 16
    a, b = ll.Context(scan_line=line), ll.Context(scan_line=line)
    ocfg = orc.config(line, voxel_stable=1)

    def wire(c):
        m = np.full((len(c), 8), 9.0, np.float32)      # padding filled with junk: it must not be read
        m[:, 0:3] = c[:, 0:3]
        m[:, 4] = c[:, 3]
        return m

    for k in range(4):
        f = orc.extract_features(ll.synth.scan(line, k), ocfg)
        pa = a.odometry_step(f["sharp"], f["less_sharp"], f["flat"], f["less_flat"])
        pb = b.odometry_step(wire(f["sharp"]), wire(f["less_sharp"]), wire(f["flat"]), wire(f["less_flat"]))
        assert np.array_equal(pa["q_w"], pb["q_w"]) and np.array_equal(pa["t_w"], pb["t_w"]), k
    a.close(); b.close()


def test_bench_configuration_parity_256_lanes(ll, orc):
    """The configuration bench.py times (HDL-64, 256 lanes, scan pool path, 256-thread solve CTAs, two per SM) against the
    oracle: 9 steps (past now_frame > 5), 8 sampled lanes incl. lanes >= 149 and lane 255.  Feature indices and association
    index triples bit-exact, per-step pose <= 1e-9 (same voxel order) and <= 1e-4 m / rad (reference-faithful order)."""
    import bench
    B, steps = 256, 9
    lanes = [0, 37, 100, 148, 149, 200, 254, 255]
    pool = bench.make_pool(ll, bench.POOL_SCANS)
    ctx = ll.Context(scan_line=64, batch=B)
    ctx.pool_upload(pool)
    ocfg = orc.config(64, voxel_stable=1)
    odos = {i: orc.Odometry(ocfg) for i in lanes}
    faithful = {i: orc.Pipeline(orc.config(64, voxel_stable=0), with_mapping=False) for i in lanes}
    for k in range(steps):
        ids = bench.lane_ids(k, B, 0)
        assert len(set(ids.tolist())) == B                      # every lane reads its own scan
        pg = ctx.process_pool(ids)
        assert not ctx.lane_status().any()
        for i in lanes:
            scan = pool[int(ids[i])]
            f = orc.extract_features(scan, ocfg)
            gf = ctx.debug_features(i)
            for key in ("sharp_idx", "less_sharp_idx", "flat_idx"):
                assert np.array_equal(gf[key], f[key]), (k, i, key)
            po = odos[i].step(f["sharp"], f["less_sharp"], f["flat"], f["less_flat"])
            assert np.abs(pg[i][4:7] - po["t_w"]).max() < 1e-9 and np.abs(pg[i][0:4] - po["q_w"]).max() < 1e-9, (k, i)
            pf = faithful[i].step(scan)
            assert np.abs(pg[i][4:7] - pf["t_odom"]).max() < 1e-4 and _ang(pg[i][0:4], pf["q_odom"]) < 1e-4, (k, i)
            if k == 0:
                continue
            oc, op = odos[i].assoc(len(f["sharp"]), len(f["flat"]))
            gc, gp = ctx.debug_assoc(i)
            gc, gp = gc[:len(f["sharp"])], gp[:len(f["flat"])]
            got_c = np.array([[q, a, b] for q, (a, b) in enumerate(gc) if b >= 0], np.int32).reshape(-1, 3)
            got_p = np.array([[q, a, b, c] for q, (a, b, c, _) in enumerate(gp) if a >= 0], np.int32).reshape(-1, 4)
            assert np.array_equal(got_c, oc), (k, i)
            assert np.array_equal(got_p, op), (k, i)
    ctx.close()


def test_packed_submit_equals_per_scan_calls(ll):
    """ll_submit_packed (one host arena of packed 12-byte xyz records, one copy per batch) == ll_process_scans on float4 scans."""
    line, B, n = 16, 3, 5
    a = ll.Context(scan_line=line, batch=B)
    b = ll.Context(scan_line=line, batch=B)
    for k in range(n):
        scans = [ll.synth.scan(line, k + 2 * i) for i in range(B)]
        pa = a.process_scans(scans)
        xyz = [np.ascontiguousarray(s[:, :3]) for s in scans]
        offs = np.concatenate([[0], np.cumsum([x.nbytes for x in xyz])]).astype(np.int64)
        arena = np.concatenate([x.reshape(-1) for x in xyz]).view(np.uint8)
        b.submit_packed(arena, offs[:-1], [len(x) for x in xyz], 12)
        pb = b.collect()
        assert np.array_equal(pa, pb), k
    a.close()
    b.close()


def test_lane_status_bad_scan_does_not_advance_the_stream(ll):
    """A scan without a valid point (the reference would index points[0] of an empty cloud, SR:114): the lane reports
    LL_E_EMPTY in its status, the call returns that code after writing every pose, and the lane's state is untouched -
    the next good scan continues as if the bad one had never arrived.  Other lanes are not affected."""
    line = 16
    ctx = ll.Context(scan_line=line, batch=2)
    ref = ll.Context(scan_line=line, batch=2)
    bad = np.full((500, 4), np.nan, np.float32)
    seq = [ll.synth.scan(line, k) for k in range(5)]
    for k in range(5):
        p = ctx.process_scans([seq[k], seq[k]])
        r = ref.process_scans([seq[k], seq[k]])
        assert np.array_equal(p, r)
        if k == 2:
            before = p.copy()
            p = ctx.process_scans([seq[3], bad])
            assert ctx.last_rc == ll.capi.LL_E_EMPTY
            assert list(ctx.lane_status()) == [0, ll.capi.LL_E_EMPTY]
            assert np.array_equal(p[1], before[1])                       # pose unchanged
            r2 = ref.process_scans([seq[3], seq[3]])
            assert np.array_equal(p[0], r2[0])                           # the good lane ran normally
            # lane 1 now catches up with the scan it missed: same result as the reference context's lane 1
            p = ctx.process_scans([seq[4], seq[3]])
            r3 = ref.process_scans([seq[4], seq[4]])
            assert list(ctx.lane_status()) == [0, 0]
            assert np.array_equal(p[1], r2[1]) and np.array_equal(p[0], r3[0])
            break
    ctx.close()
    ref.close()


@pytest.mark.parametrize("distortion,line", [(1, 16), (2, 16), (1, 64), (2, 32)])
def test_deskew_mode_matches_oracle(ll, orc, distortion, line):
    """DISTORTION 1 (LO:23, compiled out in the reference build): per-point interpolation ratio s = relTime in
    TransformToStart (LO:77-95) and in the factors (LF:23-31, 227-235; s != 1 takes the dual-number path of the solve),
    and - distortion 2 - TransformToEnd of the clouds that become *Last (LO:98-114, 861-880, behind `if (0)` in the
    reference).  Association index triples equal the oracle's, poses within 1e-9; with mapping on, the mapped pose within 1e-7."""
    ocfg = orc.config(line, voxel_stable=1, distortion=distortion)
    ctx = ll.Context(scan_line=line, distortion=distortion)
    odo = orc.Odometry(ocfg)
    plain = ll.Context(scan_line=line)
    moved = False
    for k in range(8):
        f = orc.extract_features(ll.synth.scan(line, k), ocfg)
        po = odo.step(f["sharp"], f["less_sharp"], f["flat"], f["less_flat"])
        pg = ctx.odometry_step(f["sharp"], f["less_sharp"], f["flat"], f["less_flat"])
        pp = plain.odometry_step(f["sharp"], f["less_sharp"], f["flat"], f["less_flat"])
        # distortion 2 rewrites the *Last clouds through slerp + rotate and stores them as fp32: the device's sincos / acos
        # differ from glibc's in the last bit now and then, which moves a stored coordinate by one fp32 ulp
        tol = 1e-9 if distortion == 1 else 1e-7
        assert np.abs(pg["t_w"] - po["t_w"]).max() < tol and np.abs(pg["q_w"] - po["q_w"]).max() < tol, k
        assert np.abs(pg["t_last"] - po["t_last"]).max() < tol, k
        if k == 0:
            continue
        moved |= bool(np.abs(pg["t_w"] - pp["t_w"]).max() > 1e-6)
        oc, op = odo.assoc(len(f["sharp"]), len(f["flat"]))
        gc, gp = ctx.debug_assoc(0)
        gc, gp = gc[:len(f["sharp"])], gp[:len(f["flat"])]
        got_c = np.array([[i, a, b] for i, (a, b) in enumerate(gc) if b >= 0], np.int32).reshape(-1, 3)
        got_p = np.array([[i, a, b, c] for i, (a, b, c, _) in enumerate(gp) if a >= 0], np.int32).reshape(-1, 4)
        assert np.array_equal(got_c, oc) and np.array_equal(got_p, op), k
        st, stats = ctx.stats(), odo.stats()
        assert [int(s[5]) for s in stats] == list(st.lm_jacobian_evals) and [int(s[7]) for s in stats] == list(st.lm_termination), k
    assert moved      # the mode does change the estimate
    ctx.close()
    plain.close()


def test_deskew_fused_pipeline_with_mapping(ll, orc):
    line = 16
    ctx = ll.Context(scan_line=line, distortion=2, enable_mapping=1, map_capacity=1 << 18)
    pipe = orc.Pipeline(orc.config(line, voxel_stable=1, distortion=2), with_mapping=True)
    for k in range(7):
        scan = ll.synth.scan(line, k)
        pg, pe = ctx.process_scans([scan])[0], pipe.step(scan)
        # one-ulp differences of the fp32 clouds TransformToEnd stores (device vs glibc trigonometry) reach the odometry pose
        # at ~1e-8; in the mapping stage such a coordinate can change voxel (floor(x / leaf) is discontinuous), which moves a
        # centroid and with it the mapped pose by up to ~1e-4: the bar here is the north-star one
        assert np.abs(pg[4:7] - pe["t_odom"]).max() < 1e-6 and np.abs(pg[0:4] - pe["q_odom"]).max() < 1e-6, k
        assert np.abs(pg[11:14] - pe["t_map"]).max() < 1e-3 and np.abs(pg[7:11] - pe["q_map"]).max() < 1e-3, k
    ctx.close()


@pytest.mark.parametrize("line", [16, 64])
def test_vote_partial_mode_matches_oracle(ll, orc, line):
    """vote_mode = 1: the paper-style scoring graph_based_correspondence_vote_partial (LM:321-834, dead code in the
    reference - beyond-reference behaviour, flagged as such) on the odometry's plane correspondences.  Selected counts equal
    the oracle's; poses within 1e-6 (the m x m expf / cbrt scores are fp32 roundings of fp64 evaluations on both sides, but
    not the same libm)."""
    ocfg = orc.config(line, voxel_stable=1, vote_mode=1)
    ctx = ll.Context(scan_line=line, vote_mode=1)
    simple = ll.Context(scan_line=line)
    odo = orc.Odometry(ocfg)
    differs = False
    for k in range(9):
        f = orc.extract_features(ll.synth.scan(line, k), ocfg)
        po = odo.step(f["sharp"], f["less_sharp"], f["flat"], f["less_flat"])
        pg = ctx.odometry_step(f["sharp"], f["less_sharp"], f["flat"], f["less_flat"])
        ps = simple.odometry_step(f["sharp"], f["less_sharp"], f["flat"], f["less_flat"])
        assert np.abs(pg["t_w"] - po["t_w"]).max() < 1e-6 and np.abs(pg["q_w"] - po["q_w"]).max() < 1e-6, k
        if k > 6:
            st, stats = ctx.stats(), odo.stats()
            assert [int(s[1]) for s in stats] == list(st.plane_corr), k
            assert [int(s[2]) for s in stats] == list(st.plane_selected), (k, [int(s[2]) for s in stats], list(st.plane_selected))
            assert st.plane_selected[2] < st.plane_corr[2]
            differs |= bool(np.abs(pg["t_w"] - ps["t_w"]).max() > 1e-9)
    assert differs
    ctx.close()
    simple.close()


@pytest.mark.parametrize("line,mode", [(64, 0), (64, 1), (16, 1)])
def test_long_trajectory_stays_on_the_oracle(ll, orc, line, mode):
    """40 consecutive scans: discrete decisions (nearest neighbours, votes, LM accept / reject / termination) must keep matching the
    oracle long after the vote gate, and the accumulated pose must not drift away from it.  Bar: accumulated pose within 1e-8 of the
    oracle run in the GPU's voxel order at every scan, identical correspondence / selection / evaluation counts at the end."""
    n = 40
    ctx = ll.Context(scan_line=line)
    exact = orc.Pipeline(orc.config(line, voxel_stable=1), with_mapping=False)
    worst = 0.0
    for k in range(n):
        scan = ll.synth.scan(line, k, mode=mode)
        pg = ctx.process_scans([scan])[0]
        pe = exact.step(scan)
        worst = max(worst, np.abs(pg[4:7] - pe["t_odom"]).max(), np.abs(pg[0:4] - pe["q_odom"]).max())
        assert worst < 1e-8, (k, worst)
    st = ctx.stats()
    assert st.frame == n and np.abs(pg[4:7]).max() > 1.0
    ctx.close()
