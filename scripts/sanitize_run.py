"""Small runs of every kernel family for compute-sanitizer (memcheck / racecheck): the default fused pipeline, the
TMA/SMEM-staged association (LL_ASSOC_SLAB=1 in the environment), the wide-ring kernels, vote_partial, de-skew with
TransformToEnd, scan-to-map with its graph vote, the packed asynchronous submission.  usage: python scripts/sanitize_run.py [mode]"""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ll = importlib.import_module("light-loam_b200")
mode = sys.argv[1] if len(sys.argv) > 1 else "default"
steps = int(os.environ.get("LL_STEPS", "8"))
if mode == "default":            # HDL-64, 2 lanes, past the vote gate; packed submission on the way
    ctx = ll.Context(scan_line=64, batch=2)
    for k in range(steps):
        ctx.process_scans([ll.synth.scan(64, k), ll.synth.scan(64, k + 3)])
    xyz = [np.ascontiguousarray(ll.synth.scan(64, steps + i)[:, :3]) for i in range(2)]
    offs = np.concatenate([[0], np.cumsum([x.nbytes for x in xyz])]).astype(np.int64)
    ctx.submit_packed(np.concatenate([x.reshape(-1) for x in xyz]).view(np.uint8), offs[:-1], [len(x) for x in xyz], 12)
    print("poses", ctx.collect()[:, 4:7])
elif mode == "modes":            # 16 lines: wide rings, vote_partial, de-skew + mapping + map vote
    c1 = ll.Context(scan_line=16, max_points=65536)
    c1.extract_features(ll.synth.scan(16, 0, az_steps=4000))
    c2 = ll.Context(scan_line=16, vote_mode=1, distortion=2, enable_mapping=1, map_capacity=1 << 17, map_graph_vote=1)
    for k in range(steps):
        p = c2.process_scans([ll.synth.scan(16, k)])
    print("pose", p[0][4:7], p[0][11:14])
print("done")
