"""Independent re-derivations of the order-dependent parts of the reference, written from the reference sources
(scanRegistration.cpp = SR, laserOdometry.cpp = LO) in plain Python / NumPy and sharing no code with oracle/: they pin
the oracle where the reference itself ships no tests.  CPU only."""
import numpy as np
import pytest

f32 = np.float32


def _ring_and_reltime(points, line, min_range, lower=-24.9, upper=2.0):
    """SR:58-85 (range filter), SR:113-210 (startOri / endOri, ring id, halfPassed state machine, relTime)."""
    p = np.asarray(points, np.float32)[:, :3]
    ok = np.isfinite(p).all(axis=1)
    p = p[ok]
    d2 = p[:, 0] * p[:, 0] + p[:, 1] * p[:, 1] + p[:, 2] * p[:, 2]
    p = p[~(d2 < f32(min_range) * f32(min_range))]
    n = len(p)
    start = f32(-np.arctan2(p[0, 1], p[0, 0]))
    end = f32(np.float64(f32(-np.arctan2(p[-1, 1], p[-1, 0]))) + 2 * np.pi)
    if np.float64(end - start) > 3 * np.pi:
        end = f32(np.float64(end) - 2 * np.pi)
    elif np.float64(end - start) < np.pi:
        end = f32(np.float64(end) + 2 * np.pi)
    factor = f32((line - 1) / (f32(upper) - f32(lower)))
    rings = [[] for _ in range(line)]
    half = False
    for i in range(n):
        x, y, z = p[i]
        angle = f32(np.float64(f32(np.arctan(z / np.sqrt(x * x + y * y)) * f32(180.0))) / np.pi)
        if line == 16:
            sid = int(np.float64((angle + f32(15.0)) / f32(2.0)) + 0.5)
            bad = sid > 15 or sid < 0
        elif line == 32:
            sid = int((np.float64(angle) + 92.0 / 3.0) * 3.0 / 4.0)
            bad = sid > 31 or sid < 0
        else:
            sid = int(np.float64((angle - f32(lower)) * factor) + 0.5)
            bad = sid >= 64 or sid < 0
        if bad:
            continue
        ori = f32(-np.arctan2(y, x))
        if not half:
            if np.float64(ori) < np.float64(start) - np.pi / 2:
                ori = f32(np.float64(ori) + 2 * np.pi)
            elif np.float64(ori) > np.float64(start) + np.pi * 3 / 2:
                ori = f32(np.float64(ori) - 2 * np.pi)
            if np.float64(ori - start) > np.pi:
                half = True
        else:
            ori = f32(np.float64(ori) + 2 * np.pi)
            if np.float64(ori) < np.float64(end) - np.pi * 3 / 2:
                ori = f32(np.float64(ori) + 2 * np.pi)
            elif np.float64(ori) > np.float64(end) + np.pi / 2:
                ori = f32(np.float64(ori) - 2 * np.pi)
        rel = f32((ori - start) / (end - start))
        rings[sid].append((x, y, z, f32(np.float64(sid) + 0.1 * np.float64(rel))))
    return rings


@pytest.mark.parametrize("line,k", [(16, 1), (32, 2)])
def test_ring_assignment_and_reltime_match_independent_python(ll, orc, line, k):
    scan = ll.synth.scan(line, k, az_steps=400 if line == 32 else None)
    cfg = orc.config(line)
    o = orc.extract_features(scan, cfg)
    rings = _ring_and_reltime(scan, line, 0.3)
    rb = o["ring_begin"]
    mism = 0
    for r in range(line):
        mine = np.array(rings[r], np.float32).reshape(-1, 4)
        theirs = o["full"][rb[r]:rb[r + 1]]
        if len(mine) != len(theirs):
            mism += abs(len(mine) - len(theirs))     # a point exactly on a ring rounding boundary (numpy vs glibc atanf)
            continue
        assert np.array_equal(mine[:, :3], theirs[:, :3]), r        # SR:209 push_back order inside a ring
        assert np.abs(mine[:, 3] - theirs[:, 3]).max() < 1e-5, r    # scanID + 0.1 * relTime
    assert mism <= 2


def _pick(full, ring_begin):
    """SR:224-368 on a ring-sorted cloud: curvature (float32, left-to-right sums), per-sector sort, greedy pick."""
    P = full[:, :3].astype(np.float32)
    n = len(P)
    curv = np.zeros(n, np.float32)
    for i in range(5, n - 5):
        d = P[i - 5].copy()
        for q in (-4, -3, -2, -1):
            d = d + P[i + q]
        d = d - f32(10) * P[i]
        for q in (1, 2, 3, 4, 5):
            d = d + P[i + q]
        curv[i] = d[0] * d[0] + d[1] * d[1] + d[2] * d[2]
    picked = np.zeros(n + 8, np.int32)
    label = np.zeros(n, np.int32)
    sharp, less_sharp, flat = [], [], []

    def gap2(a, b):
        d = P[a] - P[b]
        return np.float64(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])

    def suppress(ind):
        picked[ind] = 1
        for l in range(1, 6):
            if gap2(ind + l, ind + l - 1) > 0.05:
                break
            picked[ind + l] = 1
        for l in range(-1, -6, -1):
            if gap2(ind + l, ind + l + 1) > 0.05:
                break
            picked[ind + l] = 1

    for r in range(len(ring_begin) - 1):
        s, e = ring_begin[r] + 5, ring_begin[r + 1] - 6
        if e - s < 6:
            continue
        for j in range(6):
            sp, ep = s + (e - s) * j // 6, s + (e - s) * (j + 1) // 6 - 1
            order = sorted(range(sp, ep + 1), key=lambda i: (curv[i], i))   # comp = by curvature; no ties on these scans
            npk = 0
            for ind in reversed(order):
                if picked[ind] == 0 and np.float64(curv[ind]) > 0.1:
                    npk += 1
                    if npk <= 2:
                        label[ind] = 2
                        sharp.append(ind)
                        less_sharp.append(ind)
                    elif npk <= 20:
                        label[ind] = 1
                        less_sharp.append(ind)
                    else:
                        break
                    suppress(ind)
            nsm = 0
            for ind in order:
                if picked[ind] == 0 and np.float64(curv[ind]) < 0.1:
                    label[ind] = -1
                    flat.append(ind)
                    nsm += 1
                    if nsm >= 4:
                        break
                    suppress(ind)
    return curv, label, sharp, less_sharp, flat


def test_greedy_pick_matches_independent_python(ll, orc):
    line = 16
    o = orc.extract_features(ll.synth.scan(line, 4), orc.config(line))
    assert o["sort_ties"] == 0
    curv, label, sharp, less_sharp, flat = _pick(o["full"], o["ring_begin"])
    n = len(curv)
    assert np.array_equal(curv[5:n - 5], o["curvature"][5:n - 5])
    assert sharp == list(o["sharp_idx"]) and less_sharp == list(o["less_sharp_idx"]) and flat == list(o["flat_idx"])
    assert np.array_equal(label, o["label"])


def _assoc(queries, last, plane):
    """LO:491-556 (corners) / LO:653-723 (planes) with the identity transform; brute-force 1-NN in float32."""
    out = []
    L = last[:, :3].astype(np.float32)
    ring = last[:, 3].astype(np.int32)          # int(intensity)
    for i, q in enumerate(queries[:, :3].astype(np.float32)):
        d = L - q
        d2 = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1] + d[:, 2] * d[:, 2]
        c = int(np.argmin(d2))                   # ties: lowest index
        if not d2[c] < 25.0:
            continue
        rc = ring[c]
        m2, m3, i2, i3 = 25.0, 25.0, -1, -1
        for j in range(c + 1, len(L)):
            if ring[j] > rc + 2.5:
                break
            if plane:
                if ring[j] <= rc and d2[j] < m2:
                    m2, i2 = d2[j], j
                elif ring[j] > rc and d2[j] < m3:
                    m3, i3 = d2[j], j
            else:
                if ring[j] <= rc:
                    continue
                if d2[j] < m2:
                    m2, i2 = d2[j], j
        for j in range(c - 1, -1, -1):
            if ring[j] < rc - 2.5:
                break
            if plane:
                if ring[j] >= rc and d2[j] < m2:
                    m2, i2 = d2[j], j
                elif ring[j] < rc and d2[j] < m3:
                    m3, i3 = d2[j], j
            else:
                if ring[j] >= rc:
                    continue
                if d2[j] < m2:
                    m2, i2 = d2[j], j
        if plane and i2 >= 0 and i3 >= 0:
            out.append([i, c, i2, i3])
        if not plane and i2 >= 0:
            out.append([i, c, i2])
    return np.array(out, np.int32).reshape(-1, 4 if plane else 3)


def test_odometry_association_matches_independent_bruteforce(ll, orc):
    """Queries that ARE points of the last clouds: every residual is exactly zero, so the three solves leave the pose at
    the identity and the correspondences of the last outer iteration must be those of the literal loops with
    pointSel = the point itself (closest point = the point, 2nd / 3rd point from the ring window)."""
    line = 16
    cfg = orc.config(line, voxel_stable=1)
    f = orc.extract_features(ll.synth.scan(line, 2), cfg)
    ls, lf = f["less_sharp"], f["less_flat"]
    sharp, flat = ls[3::9][:180].copy(), lf[5::17][:380].copy()
    odo = orc.Odometry(cfg)
    odo.step(sharp, ls, flat, lf)
    po = odo.step(sharp, ls, flat, lf)
    assert np.abs(po["t_last"]).max() < 1e-12 and np.abs(po["q_last"][:3]).max() < 1e-12
    oc, op = odo.assoc(len(sharp), len(flat))
    mc = _assoc(sharp, ls, plane=False)
    mp = _assoc(flat, lf, plane=True)
    assert len(mc) > 50 and len(mp) > 100
    assert np.array_equal(mc, oc)
    assert np.array_equal(mp, op)
