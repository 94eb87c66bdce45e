"""ctypes binding of host/libll_synth.so — the seeded synthetic scan generator (SURVEY.md §8d scene S)."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

AZ_STEPS = {64: 2031, 32: 2170, 16: 1000}
SEED = 20240919


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(os.path.join(_HERE, "host", "libll_synth.so"))
        L.ll_synth_scan.restype = ctypes.c_int
        L.ll_synth_scan.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.POINTER(ctypes.c_double),
                                    ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_int]
        L.ll_synth_pose.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
        _LIB = L
    return _LIB


def pose(k, mode=0, seed=SEED):
    out = (ctypes.c_double * 4)()
    lib().ll_synth_pose(mode, seed, k, out)
    return np.array(out)


def scan(scan_line=64, k=0, mode=0, seed=SEED, az_steps=None, noise=0.02, lower_bound=-24.9, up_bound=2.0, scan_id=None):
    """Returns (n, 4) float32 x,y,z,0 of scan k along the canonical path `mode` (0: 1 m + 0.01 rad per scan, 1: 25 m loop)."""
    az = az_steps or AZ_STEPS[scan_line]
    p = (ctypes.c_double * 4)(*pose(k, mode, seed))
    buf = np.zeros((scan_line * az, 4), np.float32)
    n = lib().ll_synth_scan(scan_line, az, seed, k if scan_id is None else scan_id, p, lower_bound, up_bound, noise, buf.ctypes.data, buf.shape[0])
    if n < 0:
        raise ValueError("ll_synth_scan: bad arguments")
    return buf[:n].copy()
