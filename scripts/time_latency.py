"""Single-stream latency: one context, one lane, every call synchronous (what a ROS node sees per scan).  The scan arrives in
a pageable host buffer (a message payload as the middleware hands it over) or in a pinned one (a node that allocates its
message pool with cudaHostAlloc / registers it once).
usage: python scripts/time_latency.py"""
import importlib, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ll = importlib.import_module("light-loam_b200")
scans = [ll.synth.scan(64, k, mode=1) for k in range(40)]
pinned = []
for s in scans:
    t = torch.empty(s.shape, dtype=torch.float32, pin_memory=True)
    t.numpy()[...] = s
    pinned.append(t)
for mapping in (0, 1):
    for kind, src in (("pageable", scans), ("pinned", [t.numpy() for t in pinned])):
        ctx = ll.Context(scan_line=64, batch=1, enable_mapping=mapping, map_capacity=1 << 20)
        for k in range(10):
            ctx.process_scans([src[k]])
        t0 = time.perf_counter()
        for k in range(10, 40):
            p = ctx.process_scans([src[k]])
        dt = (time.perf_counter() - t0) / 30
        print("mapping=%d %s host buffer: %.3f ms per scan (buffer in, pose out, synchronous), %d launches" % (mapping, kind, dt * 1e3, ctx.stats().kernel_launches))
        ctx.close()
