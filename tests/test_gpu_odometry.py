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


@pytest.mark.parametrize("line,n,az", [(16, 10, None), (64, 9, None), (32, 8, 1200)])
def test_fused_pipeline_trajectory_matches_oracle(ll, orc, line, n, az):
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
