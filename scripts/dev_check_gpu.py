# first GPU check: feature parity vs oracle on one HDL-64 scan + a 12-scan odometry run
import importlib, sys, time, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
ll = importlib.import_module('light-loam_b200'); import orc_py as O
for line in (64, 16, 32):
    a = ll.synth.scan(line, 3)
    ctx = ll.Context(scan_line=line)
    t=time.time(); g = ctx.extract_features(a); t1=time.time()-t
    o = O.extract_features(a, O.config(line, voxel_stable=1))
    print(line, 'n_full', len(g['full']), len(o['full']), 'time', round(t1*1e3,2))
    for k in ('full','ring_begin','curvature','sharp_idx','less_sharp_idx','flat_idx','less_flat'):
        same = g[k].shape == o[k].shape and np.array_equal(g[k], o[k])
        extra = ''
        if not same and g[k].shape == o[k].shape:
            d = np.abs(g[k].astype(np.float64)-o[k].astype(np.float64)); extra = 'maxdiff %g nbad %d percol %s' % (d.max(), (d>0).sum(), (d>0).sum(axis=0) if d.ndim==2 else '')
        print('  ', k, g[k].shape, o[k].shape, 'EXACT' if same else 'DIFF', extra)
    ctx.close()
# odometry
ctx = ll.Context(scan_line=64)
P = O.Pipeline(O.config(64, voxel_stable=1), False)
for k in range(10):
    a = ll.synth.scan(64, k)
    t=time.time(); pg = ctx.process_scans([a])[0]; tg=time.time()-t
    po = P.step(a)
    s = ctx.stats()
    print(k, 'gpu t', np.round(pg[4:7],5), 'orc t', np.round(po['t_odom'],5), 'dq', np.abs(pg[0:4]-po['q_odom']).max(), 'dt', np.abs(pg[4:7]-po['t_odom']).max(), 'ms', round(tg*1e3,2), list(s.corner_corr), list(s.plane_corr), list(s.plane_selected), list(s.lm_jacobian_evals), ctx.last_timings())
