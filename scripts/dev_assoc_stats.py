import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ll = importlib.import_module("light-loam_b200")
os.environ["LL_DEBUG_ASSOC"] = "1"
for shells in (1, 2, 3):
    os.environ["LL_PLANE_SHELLS"] = str(shells)
    ctx = ll.Context(scan_line=64)
    for k in range(10):
        ctx.process_scans([ll.synth.scan(64, k, mode=1)])
    print("shells", shells, "(cumulative over 9 frames, outer iteration 2)")
    st = ctx.stats()
    print("  plane_corr", list(st.plane_corr), "n_flat", st.n_flat)
    ctx.close()
