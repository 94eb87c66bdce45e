"""ORACLE (test infrastructure) — ctypes bindings of oracle/liborc.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
The product package (light-loam_b200/) never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OrcConfig(ctypes.Structure):
    _fields_ = [("scan_line", ctypes.c_int), ("minimum_range", ctypes.c_float), ("lower_bound", ctypes.c_float),
                ("up_bound", ctypes.c_float), ("line_res", ctypes.c_float), ("plane_res", ctypes.c_float),
                ("skip_frame", ctypes.c_int), ("voxel_stable", ctypes.c_int), ("graph_from_frame", ctypes.c_int),
                ("map_graph_vote", ctypes.c_int), ("distortion", ctypes.c_int), ("vote_mode", ctypes.c_int)]


def config(scan_line=64, minimum_range=None, lower_bound=-24.9, up_bound=2.0, line_res=None, plane_res=None,
           voxel_stable=0, graph_from_frame=5, map_graph_vote=0, distortion=0, vote_mode=0):
    """Launch-file values: HDL-64 -> 5 m / 0.4 / 0.8 (launch/aloam_velodyne_HDL_64.launch:2-12),
    VLP-16 and HDL-32 -> 0.3 m / 0.2 / 0.4 (launch/aloam_velodyne_VLP_16.launch:3-13)."""
    if minimum_range is None:
        minimum_range = 5.0 if scan_line == 64 else 0.3
    if line_res is None:
        line_res = 0.4 if scan_line == 64 else 0.2
    if plane_res is None:
        plane_res = 0.8 if scan_line == 64 else 0.4
    return OrcConfig(scan_line, minimum_range, lower_bound, up_bound, line_res, plane_res, 1, voxel_stable, graph_from_frame,
                     map_graph_vote, distortion, vote_mode)


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liborc.so"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liborc.so")
        if not os.path.exists(path):
            build()
        L = ctypes.CDLL(path)
        L.orc_features_run.restype = ctypes.c_void_p
        L.orc_odom_create.restype = ctypes.c_void_p
        L.orc_map_create.restype = ctypes.c_void_p
        L.orc_pipeline_create.restype = ctypes.c_void_p
        L.orc_map_total_points.restype = ctypes.c_longlong
        _LIB = L
    return _LIB


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a):
    return ctypes.c_void_p(a.ctypes.data) if a is not None else None


def extract_features(points, cfg):
    """points: (n, >=3) float32. Returns dict of arrays (scanRegistration.cpp:87-428)."""
    L = lib()
    pts = _f32(points)
    rc = ctypes.c_int(0)
    h = ctypes.c_void_p(L.orc_features_run(_p(pts), pts.shape[0], pts.shape[1], ctypes.byref(cfg), ctypes.byref(rc)))
    try:
        if rc.value != 0:
            raise RuntimeError("orc extract_features rc=%d" % rc.value)
        sz = (ctypes.c_longlong * 7)()
        L.orc_features_sizes(h, sz)
        n, ns, nls, nf, nlf, nr, ties = [int(v) for v in sz]
        out = dict(full=np.zeros((n, 4), np.float32), ring_begin=np.zeros(nr + 1, np.int32), curvature=np.zeros(n, np.float32),
                   label=np.zeros(n, np.int32), sharp_idx=np.zeros(ns, np.int32), less_sharp_idx=np.zeros(nls, np.int32),
                   flat_idx=np.zeros(nf, np.int32), less_flat=np.zeros((nlf, 4), np.float32),
                   less_flat_ring_count=np.zeros(nr, np.int32), sort_ties=ties)
        L.orc_features_copy(h, _p(out["full"]), _p(out["ring_begin"]), _p(out["curvature"]), _p(out["label"]), _p(out["sharp_idx"]),
                            _p(out["less_sharp_idx"]), _p(out["flat_idx"]), _p(out["less_flat"]), _p(out["less_flat_ring_count"]))
        out["sharp"] = out["full"][out["sharp_idx"]]
        out["less_sharp"] = out["full"][out["less_sharp_idx"]]
        out["flat"] = out["full"][out["flat_idx"]]
        return out
    finally:
        L.orc_features_free(h)


def voxel_grid(cloud, leaf, stable=False):
    L = lib()
    c = _f32(cloud)
    out = np.zeros((max(c.shape[0], 1), 4), np.float32)
    n = L.orc_voxel_grid(_p(c), c.shape[0], ctypes.c_float(leaf), int(stable), _p(out), out.shape[0])
    assert n >= 0
    return out[:n].copy()


def knn(cloud, queries, k):
    L = lib()
    c = _f32(cloud)
    q = _f32(queries)[:, :3].copy()
    idx = np.zeros((q.shape[0], k), np.int32)
    d2 = np.zeros((q.shape[0], k), np.float32)
    L.orc_knn(_p(c), c.shape[0], _p(q), q.shape[0], k, _p(idx), _p(d2))
    return idx, d2


def graph_vote(src, tgt, corner_case=False):
    L = lib()
    s, t = _f32(src), _f32(tgt)
    n = s.shape[0]
    votes = np.zeros(n, np.float32)
    sel = np.zeros(max(n, 1), np.int32)
    w = np.zeros(max(n, 1), np.float32)
    m = L.orc_graph_vote(_p(s), _p(t), n, int(corner_case), _p(votes), _p(sel), _p(w))
    return votes, sel[:m].copy(), w[:m].copy()


def evaluate(blocks, x, autodiff=True, want_jacobian=True):
    L = lib()
    b = np.ascontiguousarray(blocks, np.float64)
    xx = np.ascontiguousarray(x, np.float64)
    rows = int(np.sum(np.where(b[:, 0] == 0, 3, 1)))
    cost = ctypes.c_double(0)
    res = np.zeros(rows)
    g = np.zeros(6)
    J = np.zeros((rows, 6))
    L.orc_evaluate(_p(b), b.shape[0], _p(xx), int(autodiff), ctypes.byref(cost), _p(res), _p(g) if want_jacobian else None,
                   _p(J) if want_jacobian else None)
    return cost.value, res, g, J


def manifold_plus(x, delta):
    L = lib()
    out = np.zeros(7)
    L.orc_manifold_plus(_p(np.ascontiguousarray(x, np.float64)), _p(np.ascontiguousarray(delta, np.float64)), _p(out))
    return out


def solve(blocks, x, max_iters=4, autodiff=True):
    L = lib()
    b = np.ascontiguousarray(blocks, np.float64)
    xx = np.array(x, np.float64)
    summ = np.zeros(6)
    iters = np.zeros((8, 8))
    L.orc_solve(_p(b), b.shape[0], _p(xx), max_iters, int(autodiff), _p(summ), _p(iters))
    return xx, summ, iters[: int(summ[2])]


def sym_eig3(A):
    L = lib()
    a = np.ascontiguousarray(A, np.float64)
    ev = np.zeros(3)
    evec = np.zeros((3, 3))
    L.orc_sym_eig3(_p(a), _p(ev), _p(evec))
    return ev, evec


def plane_fit5(pts):
    L = lib()
    p = np.ascontiguousarray(pts, np.float64)
    n = np.zeros(3)
    ok = L.orc_plane_fit5(_p(p), _p(n))
    return n, bool(ok)


class Odometry:
    def __init__(self, cfg):
        self.L = lib()
        self.h = ctypes.c_void_p(self.L.orc_odom_create(ctypes.byref(cfg)))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_odom_destroy(self.h)
            self.h = None

    def step(self, sharp, less_sharp, flat, less_flat):
        a, b, c, d = _f32(sharp), _f32(less_sharp), _f32(flat), _f32(less_flat)
        pose = np.zeros(14)
        self.L.orc_odom_step(self.h, _p(a), a.shape[0], _p(b), b.shape[0], _p(c), c.shape[0], _p(d), d.shape[0], _p(pose))
        return dict(q_w=pose[0:4].copy(), t_w=pose[4:7].copy(), q_last=pose[7:11].copy(), t_last=pose[11:14].copy())

    def stats(self):
        out = np.zeros((3, 8))
        n = self.L.orc_odom_stats(self.h, _p(out))
        return out[:n]

    def assoc(self, n_sharp, n_flat):
        nc, npl = ctypes.c_int(0), ctypes.c_int(0)
        c = np.zeros((max(n_sharp, 1), 3), np.int32)
        p = np.zeros((max(n_flat, 1), 4), np.int32)
        self.L.orc_odom_assoc(self.h, ctypes.byref(nc), _p(c), ctypes.byref(npl), _p(p))
        return c[: nc.value].copy(), p[: npl.value].copy()


class Mapping:
    def __init__(self, cfg):
        self.L = lib()
        self.h = ctypes.c_void_p(self.L.orc_map_create(ctypes.byref(cfg)))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_map_destroy(self.h)
            self.h = None

    def insert(self, corner, surf):
        c, s = _f32(corner), _f32(surf)
        self.L.orc_map_insert(self.h, _p(c), c.shape[0], _p(s), s.shape[0])

    def step(self, corner_last, surf_last, q_wodom, t_wodom):
        c, s = _f32(corner_last), _f32(surf_last)
        q = np.ascontiguousarray(q_wodom, np.float64)
        t = np.ascontiguousarray(t_wodom, np.float64)
        pose = np.zeros(7)
        info = np.zeros(8, np.int32)
        self.L.orc_map_step(self.h, _p(c), c.shape[0], _p(s), s.shape[0], _p(q), _p(t), _p(pose), _p(info))
        return dict(q=pose[:4].copy(), t=pose[4:].copy(), info=info)

    def total_points(self, which):
        return int(self.L.orc_map_total_points(self.h, which))


class Pipeline:
    """scanRegistration -> laserOdometry (-> laserMapping) on raw scans, timed per stage."""

    def __init__(self, cfg, with_mapping=True):
        self.L = lib()
        self.h = ctypes.c_void_p(self.L.orc_pipeline_create(ctypes.byref(cfg), int(with_mapping)))

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_pipeline_destroy(self.h)
            self.h = None

    def step(self, points):
        pts = _f32(points)
        poses = np.zeros(14)
        ms = np.zeros(3)
        counts = np.zeros(5, np.int32)
        rc = self.L.orc_pipeline_step(self.h, _p(pts), pts.shape[0], pts.shape[1], _p(poses), _p(ms), _p(counts))
        if rc != 0:
            raise RuntimeError("orc pipeline rc=%d" % rc)
        return dict(q_odom=poses[0:4].copy(), t_odom=poses[4:7].copy(), q_map=poses[7:11].copy(), t_map=poses[11:14].copy(),
                    ms=ms, counts=counts)
