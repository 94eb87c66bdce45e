"""Generates tests/golden/*.npz from the CPU oracle on seeded synthetic scans.

PARITY UNPINNED: the reference has no tests / golden vectors and cannot be compiled or imported here
(ROS + PCL + Ceres + Eigen missing), so these fixtures pin the ORACLE's behaviour (regression) — they are not
reference outputs.  Re-run after any deliberate oracle change:  python tests/golden/make_golden.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import orc_py  # noqa: E402

ll = importlib.import_module("light-loam_b200")
OUT = os.path.dirname(os.path.abspath(__file__))


def features(scan_line, k, az=None):
    scan = ll.synth.scan(scan_line, k, az_steps=az)
    f = orc_py.extract_features(scan, orc_py.config(scan_line, voxel_stable=1))
    return dict(n_in=len(scan), n_full=len(f["full"]), ring_begin=f["ring_begin"], sharp_idx=f["sharp_idx"], less_sharp_idx=f["less_sharp_idx"],
                flat_idx=f["flat_idx"], n_less_flat=len(f["less_flat"]), less_flat_head=f["less_flat"][:64],
                curvature_sum=np.float64(f["curvature"].astype(np.float64).sum()), sort_ties=f["sort_ties"])


def trajectory(scan_line, n, az=None, mapping=True):
    pipe = orc_py.Pipeline(orc_py.config(scan_line, voxel_stable=1), with_mapping=mapping)
    out = []
    for k in range(n):
        r = pipe.step(ll.synth.scan(scan_line, k, az_steps=az))
        out.append(np.concatenate([r["q_odom"], r["t_odom"], r["q_map"], r["t_map"]]))
    return np.array(out)


if __name__ == "__main__":
    orc_py.build()
    np.savez(os.path.join(OUT, "features_vlp16_k0.npz"), **features(16, 0))
    np.savez(os.path.join(OUT, "features_vlp16_k3.npz"), **features(16, 3))
    np.savez(os.path.join(OUT, "features_hdl32_k1_az600.npz"), **features(32, 1, az=600))
    np.savez(os.path.join(OUT, "features_hdl64_k2_az500.npz"), **features(64, 2, az=500))
    np.savez(os.path.join(OUT, "trajectory_vlp16_10.npz"), poses=trajectory(16, 10))
    np.savez(os.path.join(OUT, "trajectory_hdl64_az500_9.npz"), poses=trajectory(64, 9, az=500))
    print("golden fixtures written to", OUT)
