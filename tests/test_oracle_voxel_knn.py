"""Oracle voxel grid (pcl::VoxelGrid restatement) and exact k-NN (pcl::KdTreeFLANN restatement) against
independent NumPy / SciPy implementations that share no code with the oracle."""
import numpy as np
from scipy.spatial import cKDTree


def _brute_voxel(cloud, leaf):
    inv = np.float32(1.0) / np.float32(leaf)
    mn = cloud[:, :3].min(0)
    mx = cloud[:, :3].max(0)
    min_b = np.floor(mn * inv).astype(np.int64)
    max_b = np.floor(mx * inv).astype(np.int64)
    div = max_b - min_b + 1
    ijk = (np.floor(cloud[:, :3] * inv) - min_b.astype(np.float32)).astype(np.int64)
    idx = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    out = []
    for v in np.unique(idx):
        sel = cloud[idx == v]
        s = np.zeros(4, np.float32)
        for p in sel:           # stable input order, fp32 running sums
            s = s + p
        out.append(s / np.float32(len(sel)))
    return np.array(out, np.float32)


def test_voxel_grid_stable_matches_bruteforce(orc):
    rng = np.random.default_rng(3)
    cloud = (rng.normal(size=(3000, 4)) * [6, 4, 1, 0.01] + [0, 0, 0, 7]).astype(np.float32)
    got = orc.voxel_grid(cloud, 0.4, stable=True)
    want = _brute_voxel(cloud, 0.4)
    assert got.shape == want.shape and np.array_equal(got, want)


def test_voxel_grid_std_sort_mode_differs_only_in_last_bits(orc):
    rng = np.random.default_rng(4)
    cloud = (rng.normal(size=(5000, 4)) * [3, 3, 0.5, 0.01] + [0, 0, 0, 3]).astype(np.float32)
    a = orc.voxel_grid(cloud, 0.8, stable=True)
    b = orc.voxel_grid(cloud, 0.8, stable=False)   # PCL's std::sort: within-voxel order unspecified
    assert a.shape == b.shape
    assert np.abs(a - b).max() < 1e-5


def test_voxel_grid_edge_cases(orc):
    one = np.array([[1.0, 2.0, 3.0, 4.0]], np.float32)
    assert np.array_equal(orc.voxel_grid(one, 0.2), one)
    assert orc.voxel_grid(np.zeros((0, 4), np.float32), 0.2).shape == (0, 4)
    # leaf too small for the extent -> PCL returns the input unchanged
    far = np.array([[0, 0, 0, 0], [1e4, 1e4, 1e4, 1]], np.float32)
    assert np.array_equal(orc.voxel_grid(far, 0.001), far)


def _brute_knn(cloud, queries, k):
    idxs, d2s = [], []
    for q in queries:
        dx = q[0] - cloud[:, 0]
        dy = q[1] - cloud[:, 1]
        dz = q[2] - cloud[:, 2]
        d2 = (dx * dx + dy * dy) + dz * dz          # fp32, L2_Simple order
        order = np.lexsort((np.arange(len(cloud)), d2))[:k]
        idxs.append(order)
        d2s.append(d2[order])
    return np.array(idxs), np.array(d2s, np.float32)


def test_knn_exact_vs_bruteforce_and_scipy(orc):
    rng = np.random.default_rng(5)
    cloud = (rng.normal(size=(4000, 4)) * [20, 20, 2, 1]).astype(np.float32)
    queries = (rng.normal(size=(300, 3)) * [20, 20, 2]).astype(np.float32)
    for k in (1, 5):
        idx, d2 = orc.knn(cloud, queries, k)
        bi, bd = _brute_knn(cloud, queries, k)
        assert np.array_equal(idx, bi) and np.array_equal(d2, bd)
        _, si = cKDTree(cloud[:, :3].astype(np.float64)).query(queries.astype(np.float64), k=k)
        si = si.reshape(len(queries), k)
        assert (np.sort(si, 1) == np.sort(idx, 1)).mean() > 0.999   # fp64 tree may flip exact fp32 near-ties


def test_knn_ties_lowest_index_and_small_clouds(orc):
    cloud = np.array([[1, 0, 0, 0], [-1, 0, 0, 0], [0, 1, 0, 0], [0, -1, 0, 0]], np.float32)
    idx, d2 = orc.knn(cloud, np.zeros((1, 3), np.float32), 5)
    assert list(idx[0][:4]) == [0, 1, 2, 3] and idx[0][4] == -1
    assert np.array_equal(d2[0][:4], np.ones(4, np.float32))
