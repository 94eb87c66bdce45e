"""ctypes binding of liblightloam_b200.so (include/lightloam_b200.h) — the same stub a maintainer of the
reference would write over the C ABI.  Fails loudly when the CUDA library is missing: there is no CPU path.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LL_LIB_PATH") or os.path.join(_HERE, "liblightloam_b200.so")   # LL_LIB_PATH: development builds (statistics, tuning variants)
_LIB = None

LL_OK = 0
LL_E_INVAL, LL_E_CAPACITY, LL_E_CUDA, LL_E_NCCL, LL_E_EMPTY = -1, -2, -3, -4, -5
LL_W_FEW_CORRESPONDENCES = 1

SYMBOLS = ["ll_default_config", "ll_create", "ll_destroy", "ll_strerror", "ll_last_error", "ll_get_last_stats", "ll_reset",
           "ll_extract_features", "ll_fetch_pointcloud2", "ll_odometry_step", "ll_mapping_step", "ll_map_insert", "ll_process_scans", "ll_stage_scans",
           "ll_process_staged", "ll_submit_scans", "ll_submit_packed", "ll_collect", "ll_get_lane_status", "ll_debug_features", "ll_pool_upload", "ll_process_pool", "ll_profile_enable", "ll_profile_read", "ll_last_timings",
           "ll_debug_assoc", "ll_debug_sort_scan", "ll_cuda_stream", "ll_launch_count", "ll_comm_export", "ll_comm_local_ptr", "ll_comm_attach", "ll_comm_detach", "ll_map_set_slab"]


class LLConfig(ctypes.Structure):
    _fields_ = [("scan_line", ctypes.c_int), ("minimum_range", ctypes.c_float), ("lower_bound", ctypes.c_float),
                ("up_bound", ctypes.c_float), ("line_res", ctypes.c_float), ("plane_res", ctypes.c_float),
                ("skip_frame", ctypes.c_int), ("graph_from_frame", ctypes.c_int), ("device", ctypes.c_int),
                ("batch", ctypes.c_int), ("max_points", ctypes.c_int), ("max_ring_points", ctypes.c_int),
                ("map_capacity", ctypes.c_int), ("enable_mapping", ctypes.c_int), ("map_graph_vote", ctypes.c_int),
                ("distortion", ctypes.c_int), ("vote_mode", ctypes.c_int), ("reserved", ctypes.c_int * 3)]


class LLCloudView(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("n", ctypes.c_int), ("stride_bytes", ctypes.c_int)]


class LLCloudOut(ctypes.Structure):
    _fields_ = [("xyzi", ctypes.c_void_p), ("n", ctypes.c_int), ("cap", ctypes.c_int)]


class LLStats(ctypes.Structure):
    _fields_ = [("n_full", ctypes.c_int), ("n_sharp", ctypes.c_int), ("n_less_sharp", ctypes.c_int), ("n_flat", ctypes.c_int),
                ("n_less_flat", ctypes.c_int), ("corner_corr", ctypes.c_int * 3), ("plane_corr", ctypes.c_int * 3),
                ("plane_selected", ctypes.c_int * 3), ("lm_jacobian_evals", ctypes.c_int * 3), ("lm_cost_evals", ctypes.c_int * 3),
                ("lm_termination", ctypes.c_int * 3), ("initial_cost", ctypes.c_double * 3), ("final_cost", ctypes.c_double * 3),
                ("map_corner", ctypes.c_int), ("map_surf", ctypes.c_int), ("stack_corner", ctypes.c_int), ("stack_surf", ctypes.c_int),
                ("map_corner_corr", ctypes.c_int), ("map_surf_corr", ctypes.c_int), ("map_jacobian_evals", ctypes.c_int * 2),
                ("map_termination", ctypes.c_int * 2), ("map_initial_cost", ctypes.c_double * 2), ("map_final_cost", ctypes.c_double * 2),
                ("frame", ctypes.c_int), ("kernel_launches", ctypes.c_int), ("map_vote_corr", ctypes.c_int), ("map_vote_selected", ctypes.c_int)]


class LightLoamError(RuntimeError):
    pass


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "all"])


def lib():
    """Loads the CUDA library. Raises if it is missing — the product has no other path."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise LightLoamError("liblightloam_b200.so not built (run __graft_entry__.build()); there is no CPU fallback")
        L = ctypes.CDLL(LIB_PATH)
        L.ll_strerror.restype = ctypes.c_char_p
        L.ll_last_error.restype = ctypes.c_char_p
        L.ll_last_error.argtypes = [ctypes.c_void_p]
        L.ll_cuda_stream.restype = ctypes.c_void_p
        L.ll_cuda_stream.argtypes = [ctypes.c_void_p]
        L.ll_launch_count.argtypes = [ctypes.c_void_p]
        L.ll_create.argtypes = [ctypes.POINTER(LLConfig), ctypes.POINTER(ctypes.c_void_p)]
        L.ll_destroy.argtypes = [ctypes.c_void_p]
        L.ll_reset.argtypes = [ctypes.c_void_p]
        L.ll_get_last_stats.argtypes = [ctypes.c_void_p, ctypes.POINTER(LLStats)]
        L.ll_extract_features.argtypes = [ctypes.c_void_p, LLCloudView] + [ctypes.POINTER(LLCloudOut)] * 5 + [ctypes.c_void_p] * 5
        L.ll_odometry_step.argtypes = [ctypes.c_void_p] + [LLCloudView] * 4 + [ctypes.c_void_p] * 4
        L.ll_fetch_pointcloud2.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int)]
        L.ll_mapping_step.argtypes = [ctypes.c_void_p, LLCloudView, LLCloudView] + [ctypes.c_void_p] * 4
        L.ll_map_insert.argtypes = [ctypes.c_void_p, LLCloudView, LLCloudView]
        L.ll_process_scans.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(LLCloudView), ctypes.c_void_p]
        L.ll_stage_scans.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(LLCloudView)]
        L.ll_process_staged.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
        L.ll_submit_scans.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(LLCloudView)]
        L.ll_collect.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.ll_submit_packed.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.ll_get_lane_status.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.ll_debug_features.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
        L.ll_debug_sort_scan.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.ll_pool_upload.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.POINTER(LLCloudView)]
        L.ll_process_pool.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
        L.ll_profile_enable.argtypes = [ctypes.c_void_p, ctypes.c_int]
        L.ll_profile_read.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        L.ll_last_timings.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.ll_debug_assoc.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.ll_comm_export.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
        L.ll_comm_local_ptr.restype = ctypes.c_void_p
        L.ll_comm_local_ptr.argtypes = [ctypes.c_void_p]
        L.ll_comm_attach.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
        L.ll_comm_detach.argtypes = [ctypes.c_void_p]
        L.ll_map_set_slab.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double]
        _LIB = L
    return _LIB


def default_config(scan_line=64, **overrides):
    cfg = LLConfig()
    lib().ll_default_config(ctypes.byref(cfg), scan_line)
    for k, v in overrides.items():
        setattr(cfg, k, v)
    return cfg


def _view(a):
    if a is None or len(a) == 0:
        return LLCloudView(None, 0, 16)
    assert a.ndim == 2 and a.flags["C_CONTIGUOUS"] and a.dtype in (np.float32, np.uint8)
    return LLCloudView(a.ctypes.data, a.shape[0], a.shape[1] * a.dtype.itemsize)   # uint8 rows = raw records of any point_step


class Context:
    """One ll_ctx.  Mirrors the reference's three node bodies as methods."""

    def __init__(self, cfg=None, **overrides):
        self.L = lib()
        self.cfg = cfg if cfg is not None else default_config(**overrides)
        h = ctypes.c_void_p()
        rc = self.L.ll_create(ctypes.byref(self.cfg), ctypes.byref(h))
        if rc != LL_OK:
            raise LightLoamError("ll_create: %s" % self.L.ll_strerror(rc).decode())
        self.h = h
        self.R = self.cfg.scan_line
        self.B = self.cfg.batch

    def close(self):
        if getattr(self, "h", None):
            self.L.ll_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def _check(self, rc, what):
        if rc < 0:
            raise LightLoamError("%s: %s (%s)" % (what, self.L.ll_strerror(rc).decode(), self.L.ll_last_error(self.h).decode()))
        return rc

    def _check_batch(self, rc, what):
        """Batch calls return the first failing lane's code AFTER writing every pose: keep it in last_rc, raise only for
        call-level failures (the lane codes are in lane_status())."""
        self.last_rc = rc
        if rc < 0 and rc in (LL_E_EMPTY, LL_E_CAPACITY, LL_E_NCCL) and any(self.lane_status()):
            return rc
        return self._check(rc, what)

    def lane_status(self, n=None):
        n = self.B if n is None else n
        st = np.zeros(n, np.int32)
        self._check(self.L.ll_get_lane_status(self.h, st.ctypes.data, n), "ll_get_lane_status")
        return st

    def reset(self):
        self._check(self.L.ll_reset(self.h), "ll_reset")

    def extract_features(self, points):
        """scanRegistration.cpp:100-377. points: (n, 3..8) float32. Returns a dict like oracle/orc_py.extract_features."""
        pts = np.ascontiguousarray(points) if getattr(points, "dtype", None) == np.uint8 else np.ascontiguousarray(points, dtype=np.float32)
        n, R = pts.shape[0], self.R
        full = np.zeros((n, 4), np.float32)
        sharp = np.zeros((R * 12, 4), np.float32)
        lsharp = np.zeros((R * 120, 4), np.float32)
        flat = np.zeros((R * 24, 4), np.float32)
        lflat = np.zeros((n, 4), np.float32)
        sidx = np.zeros(R * 12, np.int32)
        lsidx = np.zeros(R * 120, np.int32)
        fidx = np.zeros(R * 24, np.int32)
        curv = np.zeros(n, np.float32)
        rb = np.zeros(R + 1, np.int32)
        outs = [LLCloudOut(a.ctypes.data, 0, a.shape[0]) for a in (full, sharp, lsharp, flat, lflat)]
        rc = self.L.ll_extract_features(self.h, _view(pts), *[ctypes.byref(o) for o in outs], sidx.ctypes.data, lsidx.ctypes.data,
                                        fidx.ctypes.data, curv.ctypes.data, rb.ctypes.data)
        self._check(rc, "ll_extract_features")
        nf, ns, nls, nfl, nlf = [o.n for o in outs]
        return dict(full=full[:nf], ring_begin=rb, curvature=curv[:nf], sharp=sharp[:ns], less_sharp=lsharp[:nls], flat=flat[:nfl],
                    less_flat=lflat[:nlf], sharp_idx=sidx[:ns], less_sharp_idx=lsidx[:nls], flat_idx=fidx[:nfl])

    def fetch_pointcloud2(self, which, cap_points=None):
        """Cloud `which` (0 full, 1 sharp, 2 less_sharp, 3 flat, 4 less_flat) of the last extraction as PointCloud2 data:
        (n, 32) uint8, pcl::PointXYZI records (SR:382-410)."""
        cap = self.cfg.max_points if cap_points is None else cap_points
        buf = np.zeros((max(cap, 1), 32), np.uint8)
        n = ctypes.c_int(0)
        rc = self.L.ll_fetch_pointcloud2(self.h, which, buf.ctypes.data, cap, ctypes.byref(n))
        self._check(rc, "ll_fetch_pointcloud2")
        return buf[: n.value]

    def odometry_step(self, sharp, less_sharp, flat, less_flat):
        """laserOdometry.cpp:425-896 for one synchronized set of feature clouds (float32 (n,4))."""
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (sharp, less_sharp, flat, less_flat)]
        qw, tw, ql, tl = np.zeros(4), np.zeros(3), np.zeros(4), np.zeros(3)
        rc = self.L.ll_odometry_step(self.h, *[_view(a) for a in arrs], qw.ctypes.data, tw.ctypes.data, ql.ctypes.data, tl.ctypes.data)
        self._check(rc, "ll_odometry_step")
        return dict(q_w=qw, t_w=tw, q_last=ql, t_last=tl, rc=rc)

    def mapping_step(self, corner_last, surf_last, q_wodom, t_wodom):
        """laserMapping.cpp:1581-2168."""
        c = np.ascontiguousarray(corner_last, dtype=np.float32)
        s = np.ascontiguousarray(surf_last, dtype=np.float32)
        qi = np.ascontiguousarray(q_wodom, dtype=np.float64)
        ti = np.ascontiguousarray(t_wodom, dtype=np.float64)
        q, t = np.zeros(4), np.zeros(3)
        rc = self.L.ll_mapping_step(self.h, _view(c), _view(s), qi.ctypes.data, ti.ctypes.data, q.ctypes.data, t.ctypes.data)
        self._check(rc, "ll_mapping_step")
        return dict(q=q, t=t, rc=rc)

    def map_insert(self, corner, surf):
        c = np.ascontiguousarray(corner, dtype=np.float32)
        s = np.ascontiguousarray(surf, dtype=np.float32)
        self._check(self.L.ll_map_insert(self.h, _view(c), _view(s)), "ll_map_insert")

    def _views(self, scans):
        self._keep = [np.ascontiguousarray(s) if getattr(s, "dtype", None) == np.uint8 else np.ascontiguousarray(s, dtype=np.float32) for s in scans]
        arr = (LLCloudView * len(scans))()
        for i, a in enumerate(self._keep):
            arr[i] = _view(a)
        return arr

    def process_scans(self, scans):
        """Fused pipeline: scan i feeds lane i. Returns (n, 14) float64 poses."""
        views = self._views(scans)
        poses = np.zeros((len(scans), 14))
        self._check_batch(self.L.ll_process_scans(self.h, len(scans), views, poses.ctypes.data), "ll_process_scans")
        return poses

    def stage_scans(self, scans):
        views = self._views(scans)
        self._check(self.L.ll_stage_scans(self.h, len(scans), views), "ll_stage_scans")

    def process_staged(self, n, want_poses=True):
        poses = np.zeros((n, 14)) if want_poses else None
        self._check_batch(self.L.ll_process_staged(self.h, n, poses.ctypes.data if want_poses else None), "ll_process_staged")
        return poses

    def submit_scans(self, scans):
        """Asynchronous ll_process_scans: returns after enqueuing; at most two submissions in flight."""
        arrs = [np.ascontiguousarray(s, dtype=np.float32) for s in scans]
        views = (LLCloudView * len(arrs))()
        for i, a in enumerate(arrs):
            views[i] = _view(a)
        if not hasattr(self, "_inflight"):
            self._inflight = []
        self._check(self.L.ll_submit_scans(self.h, len(arrs), views), "ll_submit_scans")
        self._inflight.append(arrs)      # keep the host buffers alive until collected

    def make_views(self, arrays):
        """Pre-builds the ll_cloud_view records of float32 (n, k) arrays (kept alive by the caller)."""
        return [_view(a) for a in arrays]

    def submit_views(self, views):
        """ll_submit_scans on pre-built views (no per-call array conversion); buffers must outlive the collect."""
        arr = (LLCloudView * len(views))(*views)
        if not hasattr(self, "_inflight"):
            self._inflight = []
        self._check(self.L.ll_submit_scans(self.h, len(views), arr), "ll_submit_scans")
        self._inflight.append(arr)

    def submit_packed(self, arena, offsets, counts, stride_bytes):
        """ll_submit_packed: `arena` = one (pinned) uint8 / float32 host array holding every scan, offsets in bytes."""
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        cnt = np.ascontiguousarray(counts, dtype=np.int32)
        if not hasattr(self, "_inflight"):
            self._inflight = []
        self._check(self.L.ll_submit_packed(self.h, len(cnt), arena.ctypes.data, off.ctypes.data, cnt.ctypes.data, stride_bytes), "ll_submit_packed")
        self._inflight.append((arena, len(cnt)))

    def collect(self):
        poses = np.zeros((self.B, 14))
        n = self.L.ll_collect(self.h, poses.ctypes.data)
        if not getattr(self, "_inflight", None):
            self._check(n, "ll_collect")           # nothing outstanding: LL_E_INVAL
        held = self._inflight.pop(0)
        if n < 0:
            self._check_batch(n, "ll_collect")
            n = held[1] if isinstance(held, tuple) else len(held)
        return poses[:n]

    def pool_upload(self, scans):
        """Keeps the scans resident in HBM; lanes are then fed by scan id (process_pool)."""
        views = self._views(scans)
        self._check(self.L.ll_pool_upload(self.h, len(scans), views), "ll_pool_upload")
        self._keep = None

    def process_pool(self, scan_ids, want_poses=True):
        ids = np.ascontiguousarray(scan_ids, dtype=np.int32)
        poses = np.zeros((len(ids), 14)) if want_poses else None
        self._check_batch(self.L.ll_process_pool(self.h, len(ids), ids.ctypes.data, poses.ctypes.data if want_poses else None), "ll_process_pool")
        return poses

    def profile_enable(self, on=True):
        self._check(self.L.ll_profile_enable(self.h, int(on)), "ll_profile_enable")

    def profile_read(self):
        """{kernel name: (total ms, launches)} accumulated since profile_enable(True)."""
        buf = ctypes.create_string_buffer(4096)
        ms = np.zeros(64)
        cnt = np.zeros(64, np.int32)
        n = self._check(self.L.ll_profile_read(self.h, buf, 4096, ms.ctypes.data, cnt.ctypes.data, 64), "ll_profile_read")
        names = buf.value.decode().split("\n")[:n]
        return {names[i]: (float(ms[i]), int(cnt[i])) for i in range(n)}

    def last_timings(self):
        ms = np.zeros(4, np.float32)
        self._check(self.L.ll_last_timings(self.h, ms.ctypes.data), "ll_last_timings")
        return ms

    def stats(self):
        s = LLStats()
        self._check(self.L.ll_get_last_stats(self.h, ctypes.byref(s)), "ll_get_last_stats")
        return s

    def debug_features(self, lane=0):
        """Feature indices of lane `lane`'s last extraction: dict(counts, sharp_idx, less_sharp_idx, flat_idx)."""
        cnt = np.zeros(5, np.int32)
        s_ = np.zeros(self.R * 12, np.int32)
        ls = np.zeros(self.R * 120, np.int32)
        f = np.zeros(self.R * 24, np.int32)
        self._check(self.L.ll_debug_features(self.h, lane, cnt.ctypes.data, s_.ctypes.data, ls.ctypes.data, f.ctypes.data), "ll_debug_features")
        return dict(counts=cnt, sharp_idx=s_[:cnt[1]], less_sharp_idx=ls[:cnt[2]], flat_idx=f[:cnt[3]])

    def debug_sort_scan(self, keys, vals, key_bits=64, capacity=None, scan=None):
        """The map filter's radix sort / prefix sum on caller data (ll_debug_sort_scan): returns (keys, vals, scan) after the call."""
        k = np.ascontiguousarray(keys, np.uint64).copy()
        v = np.ascontiguousarray(vals, np.int32).copy()
        sc = None if scan is None else np.ascontiguousarray(scan, np.int32).copy()
        cap = max(len(k), 0 if sc is None else len(sc)) if capacity is None else capacity
        self._check(self.L.ll_debug_sort_scan(self.h, k.ctypes.data, v.ctypes.data, len(k), key_bits, cap, None if sc is None else sc.ctypes.data,
                                              0 if sc is None else len(sc)), "ll_debug_sort_scan")
        return k, v, sc

    def debug_assoc(self, lane=0):
        c = np.zeros((self.R * 12, 2), np.int32)
        p = np.zeros((self.R * 24, 4), np.int32)
        self._check(self.L.ll_debug_assoc(self.h, lane, c.ctypes.data, c.shape[0], p.ctypes.data, p.shape[0]), "ll_debug_assoc")
        return c, p

    # ---- multi-GPU scan-to-map (configs[4]): slab-sharded queries + in-kernel all-reduce over peer memory ----------
    def comm_export(self):
        """64-byte CUDA IPC handle of this context's mailbox (send it to the other ranks)."""
        buf = ctypes.create_string_buffer(64)
        self._check(self.L.ll_comm_export(self.h, buf), "ll_comm_export")
        return buf.raw

    def comm_local_ptr(self):
        return self.L.ll_comm_local_ptr(self.h)

    def comm_attach(self, rank, world, peers):
        """peers: list of `world` 64-byte handles (other processes) or of `world` ints (pointers of contexts in this process)."""
        if isinstance(peers[0], (bytes, bytearray)):
            blob = b"".join(bytes(p) for p in peers)
            assert len(blob) == 64 * world
            self._check(self.L.ll_comm_attach(self.h, rank, world, blob, 0), "ll_comm_attach")
        else:
            arr = (ctypes.c_void_p * world)(*[int(p) for p in peers])
            self._check(self.L.ll_comm_attach(self.h, rank, world, arr, 1), "ll_comm_attach")

    def comm_detach(self):
        self._check(self.L.ll_comm_detach(self.h), "ll_comm_detach")

    def map_set_slab(self, x_lo, x_hi):
        self._check(self.L.ll_map_set_slab(self.h, float(x_lo), float(x_hi)), "ll_map_set_slab")

    def cuda_stream(self):
        return self.L.ll_cuda_stream(self.h)
