"""Single-stream latency: one context, one lane, every call synchronous (what a ROS node sees per scan).
usage: python scripts/time_latency.py"""
import importlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ll = importlib.import_module("light-loam_b200")
scans = [ll.synth.scan(64, k, mode=1) for k in range(40)]
for mapping in (0, 1):
    ctx = ll.Context(scan_line=64, batch=1, enable_mapping=mapping, map_capacity=1 << 20)
    for k in range(10):
        ctx.process_scans([scans[k]])
    t0 = time.perf_counter()
    for k in range(10, 40):
        p = ctx.process_scans([scans[k]])
    dt = (time.perf_counter() - t0) / 30
    print("mapping=%d: %.3f ms per scan (host buffer in, pose out, synchronous), %d launches" % (mapping, dt * 1e3, ctx.stats().kernel_launches))
    ctx.close()
