"""Config-3 timing (BASELINE.json configs[2]): one HDL-64 scan-to-map step against a preloaded ~1e6-point local map.
Per-kernel device times come from ll_profile_enable (CUDA event pairs on the launching stream).  The map is
re-inserted before every timed step because the step's own voxel filter (LM:2155-2168) thins it."""
import importlib, json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ll = importlib.import_module("light-loam_b200")


def config3_map(n_s=1000000, n_c=100000, seed=3):
    rng = np.random.default_rng(seed)
    surf = np.zeros((n_s, 4), np.float32)
    u = rng.uniform(-1, 1, (n_s, 2))
    which = rng.integers(0, 5, n_s)
    surf[:, 0] = np.where(which == 1, 60, np.where(which == 2, -60, u[:, 0] * 60))
    surf[:, 1] = np.where(which == 3, 40, np.where(which == 4, -40, u[:, 1] * 40))
    surf[:, 2] = np.where(which == 0, -1.73, rng.uniform(-1.73, 13, n_s))
    surf[:, :3] += rng.normal(0, 0.01, (n_s, 3))
    corner = np.zeros((n_c, 4), np.float32)
    poles = rng.uniform(-55, 55, (200, 2))
    pid = rng.integers(0, 200, n_c)
    corner[:, 0], corner[:, 1] = poles[pid, 0], poles[pid, 1] * 0.7
    corner[:, 2] = rng.uniform(-1.7, 8, n_c)
    corner[:, :3] += rng.normal(0, 0.01, (n_c, 3))
    return corner, surf


if __name__ == "__main__":
    reps = int(os.environ.get("LL_REPS", "4"))
    corner, surf = config3_map()
    ctx = ll.Context(scan_line=64, map_capacity=1 << 21)
    f = ctx.extract_features(ll.synth.scan(64, 0, mode=1))
    q0 = np.array([0, 0, np.sin(np.pi / 4), np.cos(np.pi / 4)])
    t0 = np.array([25.0, 0.0, 0.0])
    out = {}
    for r in range(reps):
        ctx.reset()
        ctx.map_insert(corner, surf)
        if r == reps - 1:
            ctx.profile_enable(True)
        m = ctx.mapping_step(f["less_sharp"], f["less_flat"], q0, t0)
        st = ctx.stats()
    prof = ctx.profile_read()
    tot = sum(v[0] for v in prof.values())
    out = {"map_corner": st.map_corner, "map_surf": st.map_surf, "stack_corner": st.stack_corner, "stack_surf": st.stack_surf,
           "corner_corr": st.map_corner_corr, "surf_corr": st.map_surf_corr, "t": list(m["t"]), "total_ms": tot,
           "kernels": {k: {"ms": round(v[0], 4), "launches": v[1]} for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}}
    print(json.dumps(out))
