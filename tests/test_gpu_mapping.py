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
