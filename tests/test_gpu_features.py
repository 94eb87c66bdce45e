"""GPU parity, feature extraction: ll_extract_features (C ABI -> sm_100a kernels) vs the oracle on identical scans.
Bar: ring order, curvature and every feature index bit-exact; x,y,z bit-exact; intensity (ring + 0.1*relTime, an
fp32 atan2 chain) integer part exact and fraction within 4e-6; less-flat voxel centroids bit-exact against the
oracle's stable-order mode and within 1e-5 of the reference-faithful std::sort mode."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _compare(g, o, exact_less_flat=True):
    assert np.array_equal(g["ring_begin"], o["ring_begin"])
    assert g["full"].shape == o["full"].shape
    assert np.array_equal(g["full"][:, :3], o["full"][:, :3])
    assert np.array_equal(np.floor(g["full"][:, 3]), np.floor(o["full"][:, 3]))
    assert np.abs(g["full"][:, 3] - o["full"][:, 3]).max(initial=0) <= 4e-6
    assert np.array_equal(g["curvature"], o["curvature"])
    for key in ("sharp_idx", "less_sharp_idx", "flat_idx"):
        assert np.array_equal(g[key], o[key]), key
    assert g["less_flat"].shape == o["less_flat"].shape
    if exact_less_flat:
        assert np.array_equal(g["less_flat"][:, :3], o["less_flat"][:, :3])
    else:
        assert np.abs(g["less_flat"][:, :3] - o["less_flat"][:, :3]).max(initial=0) <= 1e-5
    assert np.abs(g["less_flat"][:, 3] - o["less_flat"][:, 3]).max(initial=0) <= 8e-6


@pytest.mark.parametrize("line,k", [(16, 0), (16, 5), (32, 1), (64, 2), (64, 7)])
def test_features_match_oracle(ll, orc, line, k):
    scan = ll.synth.scan(line, k)
    ctx = ll.Context(scan_line=line)
    g = ctx.extract_features(scan)
    _compare(g, orc.extract_features(scan, orc.config(line, voxel_stable=1)))
    _compare(g, orc.extract_features(scan, orc.config(line, voxel_stable=0)), exact_less_flat=False)
    ctx.close()


@pytest.mark.parametrize("name,line,k,az", [("features_vlp16_k0.npz", 16, 0, None), ("features_hdl32_k1_az600.npz", 32, 1, 600),
                                           ("features_hdl64_k2_az500.npz", 64, 2, 500)])
def test_features_match_committed_golden(ll, name, line, k, az):
    gold = np.load(os.path.join(GOLD, name))
    ctx = ll.Context(scan_line=line)
    g = ctx.extract_features(ll.synth.scan(line, k, az_steps=az))
    for key in ("ring_begin", "sharp_idx", "less_sharp_idx", "flat_idx"):
        assert np.array_equal(g[key], gold[key]), key
    assert len(g["less_flat"]) == int(gold["n_less_flat"])
    assert np.array_equal(g["less_flat"][:64, :3], gold["less_flat_head"][:, :3])
    assert abs(float(g["curvature"].astype(np.float64).sum()) - float(gold["curvature_sum"])) == 0.0
    ctx.close()


def test_edge_cases_nan_close_strides_and_order(ll, orc):
    scan = ll.synth.scan(16, 1)
    ctx = ll.Context(scan_line=16)
    base = ctx.extract_features(scan)
    # NaN / inf / too-close points are dropped exactly like SR:109-110
    dirty = scan.copy()
    dirty[7, 1] = np.nan
    dirty[8, 2] = -np.inf
    dirty[9, :3] = 0.05
    g = ctx.extract_features(dirty)
    _compare(g, orc.extract_features(dirty, orc.config(16, voxel_stable=1)))
    # PointCloud2 point_step 12 / 32 (xyz at offsets 0,4,8)
    for width in (3, 8):
        wide = np.zeros((len(scan), width), np.float32)
        wide[:, :3] = scan[:, :3]
        gw = ctx.extract_features(wide)
        assert np.array_equal(gw["flat_idx"], base["flat_idx"]) and np.array_equal(gw["full"], base["full"])
    # ring-major instead of azimuth-major input order: rings keep their internal order -> same ring contents
    ang = np.round((np.degrees(np.arctan2(scan[:, 2], np.hypot(scan[:, 0], scan[:, 1]))) + 15) / 2).astype(int)
    resorted = scan[np.argsort(ang, kind="stable")]
    _compare(ctx.extract_features(resorted), orc.extract_features(resorted, orc.config(16, voxel_stable=1)))
    ctx.close()


def test_degenerate_inputs(ll, orc):
    ctx = ll.Context(scan_line=16)
    # every point closer than minimum_range -> LL_E_EMPTY (the reference would index points[0] of an empty cloud)
    with pytest.raises(ll.LightLoamError):
        ctx.extract_features(np.full((500, 4), 0.01, np.float32))
    # a scan with very short rings: rings with fewer than 17 points produce no features (SR:248)
    scan = ll.synth.scan(16, 0, az_steps=14)
    g = ctx.extract_features(scan)
    o = orc.extract_features(scan, orc.config(16, voxel_stable=1))
    assert len(g["sharp_idx"]) == len(o["sharp_idx"]) == 0 and np.array_equal(g["curvature"], o["curvature"])
    scan = ll.synth.scan(16, 0, az_steps=60)
    _compare(ctx.extract_features(scan), orc.extract_features(scan, orc.config(16, voxel_stable=1)))
    # capacity: more points than max_points
    with pytest.raises(ll.LightLoamError):
        ctx.extract_features(np.ones((40000, 4), np.float32))
    ctx.close()


def test_ring_capacity_switch_and_overflow(ll, orc):
    # 3200 points per ring needs the 1024-key sector sort; a context created with the 512-key capacity must report LL_E_CAPACITY
    scan = ll.synth.scan(16, 2, az_steps=3200)
    big = ll.Context(scan_line=16, max_points=65536, max_ring_points=3300)
    _compare(big.extract_features(scan), orc.extract_features(scan, orc.config(16, voxel_stable=1)))
    big.close()
    small = ll.Context(scan_line=16, max_points=65536, max_ring_points=3083)
    with pytest.raises(ll.LightLoamError):
        small.extract_features(scan)
    small.close()


def test_repeatable_and_idempotent(ll):
    scan = ll.synth.scan(32, 3)
    ctx = ll.Context(scan_line=32)
    a = ctx.extract_features(scan)
    b = ctx.extract_features(scan)
    for key in a:
        assert np.array_equal(a[key], b[key]), key
    ctx.close()


def test_pointcloud2_wire_format_in_and_out(ll):
    """Formats at the boundary: a PointCloud2 payload (point_step 32, x@0 y@4 z@8 intensity@16) is read in
    place through the stride, and the published clouds come back packed on the device in the same pcl::PointXYZI layout."""
    line = 16
    scan = ll.synth.scan(line, 3)
    ctx = ll.Context(scan_line=line)
    ref = ctx.extract_features(scan)
    msg = np.zeros((len(scan), 8), np.float32)            # sensor_msgs/PointCloud2::data of a PointXYZI cloud
    msg[:, 0:3] = scan[:, 0:3]
    msg[:, 4] = 7.0                                       # input intensity: ignored by scanRegistration (SR:133-209)
    got = ctx.extract_features(msg)                       # stride 32 bytes
    for k in ("full", "sharp", "less_sharp", "flat", "less_flat", "sharp_idx", "flat_idx"):
        assert np.array_equal(got[k], ref[k]), k
    for which, k in enumerate(("full", "sharp", "less_sharp", "flat", "less_flat")):
        raw = ctx.fetch_pointcloud2(which)
        assert raw.shape == (len(ref[k]), 32)
        f = raw.view(np.float32).reshape(-1, 8)
        assert np.array_equal(f[:, 0:3], ref[k][:, 0:3]) and np.array_equal(f[:, 4], ref[k][:, 3]), k
        assert (f[:, 3] == 1.0).all() and not f[:, 5:8].any(), k   # data[3] = 1.0f (PCL_ADD_POINT4D), the rest of the padding zeroed
    with pytest.raises(Exception):
        ctx.fetch_pointcloud2(0, cap_points=10)             # LL_E_CAPACITY
    ctx.close()


def test_arbitrary_point_step_records(ll, orc):
    """PointCloud2 layouts pcl::fromROSMsg accepts (SR:105-106) beyond the 12..32-byte in-place range: 22-byte XYZIRT
    (not a multiple of 4) and 48-byte records are gathered to xyz by the strided copy; results equal the float4 path."""
    scan = ll.synth.scan(16, 1)
    ctx = ll.Context(scan_line=16)
    want = ctx.extract_features(scan)
    for step in (22, 48, 12, 20):
        raw = np.zeros((len(scan), step), np.uint8)
        raw[:, :12] = np.ascontiguousarray(scan[:, :3]).view(np.uint8).reshape(-1, 12)
        raw[:, 12:] = 0xAB                                      # whatever else the driver packs behind x,y,z
        got = ctx.extract_features(raw)
        for key in ("sharp_idx", "less_sharp_idx", "flat_idx"):
            assert np.array_equal(got[key], want[key]), (step, key)
        assert np.array_equal(got["full"][:, :3], want["full"][:, :3])
    ctx.close()


def test_rings_longer_than_3083_points_take_the_wide_sector_kernels(ll, orc):
    """Default ring capacity is 6155 points (real HDL-64 scanIDs that collect two lasers, VLP-16 at 5 Hz): rings above
    3083 points run the 1024-key sector variants next to the 512-key ones; feature indices stay bit-exact.  A context
    created with the small capacity rejects the same scan with LL_E_CAPACITY and leaves the lane where it was."""
    line, az = 16, 4000
    scan = ll.synth.scan(line, 0, az_steps=az)
    ocfg = orc.config(line, voxel_stable=1)
    o = orc.extract_features(scan, ocfg)
    assert np.diff(o["ring_begin"]).max() > 3083
    ctx = ll.Context(scan_line=line, max_points=65536)
    g = ctx.extract_features(scan)
    _compare(g, o)      # indices, curvature and x, y, z bit-exact; intensity fraction within the documented 4e-6
    # mixed: a normal scan through the same context still takes the fast path with identical results
    s2 = ll.synth.scan(line, 1)
    g2, o2 = ctx.extract_features(s2), orc.extract_features(s2, ocfg)
    for key in ("sharp_idx", "less_sharp_idx", "flat_idx"):
        assert np.array_equal(g2[key], o2[key]), key
    ctx.close()
    small = ll.Context(scan_line=line, max_points=65536, max_ring_points=3083)
    with pytest.raises(ll.capi.LightLoamError):
        small.extract_features(scan)
    small.close()
