"""Oracle cost functors + LM (lidarFactor.hpp / ceres::Solve restatement) against NumPy re-derivations that
share no code with it: finite differences through the manifold, a hand-rolled LM step, and known-pose recovery."""
import numpy as np


def _rot(q):
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def _blocks(rng, n=40):
    rows = []
    for i in range(n):
        cp = rng.normal(size=3) * 10
        if i % 3 == 0:
            a = cp + rng.normal(size=3) * 0.3
            b = a + rng.normal(size=3)
            rows.append([0, *cp, *a, *b, 0, 0, 0, 1.0])
        elif i % 3 == 1:
            j = cp + rng.normal(size=3) * 0.3
            l = j + rng.normal(size=3)
            m = j + rng.normal(size=3)
            rows.append([1, *cp, *j, *l, *m, [1.0, 5.0][i % 2]])
        else:
            n_ = rng.normal(size=3)
            n_ /= np.linalg.norm(n_)
            rows.append([2, *cp, *n_, 0, 0, 0, 0, 0, 0, -float(n_ @ cp) + 0.05])
    return np.array(rows, np.float64)


def _np_residuals(blocks, x):
    R, t = _rot(x[:4]), x[4:]
    out = []
    for b in blocks:
        lp = R @ b[1:4] + t
        if b[0] == 0:
            a, bb = b[4:7], b[7:10]
            out.extend(np.cross(lp - a, lp - bb) / np.linalg.norm(a - bb))
        elif b[0] == 1:
            j, l, m = b[4:7], b[7:10], b[10:13]
            nrm = np.cross(j - l, j - m)
            nrm /= np.linalg.norm(nrm)
            out.append((lp - j) @ nrm * b[13])
        else:
            out.append(b[4:7] @ lp + b[13])
    return np.array(out)


def _huber(blocks, r):
    """per-block Huber(0.1) corrected residual scaling and cost"""
    scale, cost, k = [], 0.0, 0
    for b in blocks:
        nr = 3 if b[0] == 0 else 1
        s = float(np.sum(r[k:k + nr] ** 2))
        if s > 0.01:
            cost += 0.5 * (0.2 * np.sqrt(s) - 0.01)
            scale.extend([np.sqrt(0.1 / np.sqrt(s))] * nr)
        else:
            cost += 0.5 * s
            scale.extend([1.0] * nr)
        k += nr
    return np.array(scale), cost


def test_functor_values_match_numpy(orc):
    rng = np.random.default_rng(0)
    blocks = _blocks(rng)
    q = np.array([0.01, -0.02, 0.03, 1.0])
    q /= np.linalg.norm(q)
    x = np.concatenate([q, [0.1, -0.2, 0.05]])
    cost, res, g, J = orc.evaluate(blocks, x, autodiff=True)
    raw = _np_residuals(blocks, x)
    scale, ncost = _huber(blocks, raw)
    assert np.allclose(res, raw * scale, rtol=1e-12, atol=1e-13)
    assert abs(cost - ncost) < 1e-12 * max(1, ncost)
    assert np.allclose(g, J.T @ res, rtol=1e-10, atol=1e-12)


def test_autodiff_equals_analytic_and_finite_differences(orc):
    rng = np.random.default_rng(1)
    blocks = _blocks(rng)
    q = np.array([0.05, 0.02, -0.04, 1.0])
    q /= np.linalg.norm(q)
    x = np.concatenate([q, [0.3, 0.1, -0.2]])
    _, r1, g1, J1 = orc.evaluate(blocks, x, autodiff=True)
    _, r2, g2, J2 = orc.evaluate(blocks, x, autodiff=False)   # closed form -2[R cp]x | I (what the GPU kernel uses)
    assert np.allclose(r1, r2, rtol=1e-13, atol=1e-14) and np.allclose(J1, J2, rtol=1e-10, atol=1e-11)
    # central differences of the *uncorrected* residuals through EigenQuaternionManifold::Plus
    raw0 = _np_residuals(blocks, x)
    scale, _ = _huber(blocks, raw0)
    eps = 1e-6
    Jfd = np.zeros_like(J1)
    for c in range(6):
        d = np.zeros(6)
        d[c] = eps
        rp = _np_residuals(blocks, orc.manifold_plus(x, d))
        rm = _np_residuals(blocks, orc.manifold_plus(x, -d))
        Jfd[:, c] = (rp - rm) / (2 * eps) * scale
    assert np.allclose(J1, Jfd, rtol=1e-5, atol=1e-6)


def test_manifold_plus_matches_numpy(orc):
    x = np.array([0.1, -0.2, 0.3, 0.9, 1, 2, 3.0])
    x[:4] /= np.linalg.norm(x[:4])
    d = np.array([0.01, -0.02, 0.03, 0.5, -0.5, 0.25])
    n = np.linalg.norm(d[:3])
    dq = np.concatenate([np.sin(n) / n * d[:3], [np.cos(n)]])
    ax, ay, az, aw = dq
    bx, by, bz, bw = x[:4]
    want_q = [aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz, aw * bz + az * bw + ax * by - ay * bx,
              aw * bw - ax * bx - ay * by - az * bz]
    got = orc.manifold_plus(x, d)
    assert np.allclose(got[:4], want_q, atol=1e-15) and np.allclose(got[4:], x[4:] + d[3:])
    assert np.array_equal(orc.manifold_plus(x, np.zeros(6)), x)


def test_first_lm_iteration_matches_numpy(orc):
    """One trust-region step (radius 1e4, Jacobi scaling, clamped diagonal) recomputed with numpy.lstsq."""
    rng = np.random.default_rng(2)
    blocks = _blocks(rng, 60)
    x0 = np.array([0, 0, 0, 1, 0, 0, 0.0])
    cost0, r, g, J = orc.evaluate(blocks, x0)
    scale = 1.0 / (1.0 + np.linalg.norm(J, axis=0))
    Js = J * scale
    diag = np.clip((Js ** 2).sum(0), 1e-6, 1e32)
    D = np.sqrt(diag / 1e4)
    A = np.vstack([Js, np.diag(D)])
    y = np.linalg.lstsq(A, np.concatenate([r, np.zeros(6)]), rcond=None)[0]
    delta = -y * scale
    want = orc.manifold_plus(x0, delta)
    x1, summ, iters = orc.solve(blocks, x0, max_iters=1)
    assert summ[3] == 2 and iters[1][7] == 1       # accepted: 2 Jacobian evaluations
    assert np.allclose(x1, want, rtol=1e-9, atol=1e-11)
    assert abs(iters[0][0] - cost0) < 1e-12


def test_solve_recovers_known_pose(orc):
    rng = np.random.default_rng(7)
    q = np.array([0.01, 0.015, -0.02, 1.0])
    q /= np.linalg.norm(q)
    t = np.array([0.4, -0.1, 0.05])
    R = _rot(q)
    rows = []
    for i in range(300):
        cp = rng.normal(size=3) * 15
        lp = R @ cp + t
        if i % 2 == 0:
            d = rng.normal(size=3)
            rows.append([0, *cp, *(lp + 0.3 * d), *(lp - 0.7 * d), 0, 0, 0, 1.0])
        else:
            n_ = rng.normal(size=3)
            n_ /= np.linalg.norm(n_)
            rows.append([2, *cp, *n_, 0, 0, 0, 0, 0, 0, -float(n_ @ lp)])
    x = np.array([0, 0, 0, 1, 0, 0, 0.0])
    for _ in range(3):                                # 3 Solves of <= 4 iterations like LO:439
        x, summ, _ = orc.solve(np.array(rows), x, max_iters=4)
    assert np.allclose(x[4:], t, atol=1e-6) and min(np.abs(x[:4] - q).max(), np.abs(x[:4] + q).max()) < 1e-6
    assert summ[1] < 1e-10


def test_small_dense_helpers(orc):
    rng = np.random.default_rng(9)
    for _ in range(20):
        M = rng.normal(size=(5, 3))
        cov = (M - M.mean(0)).T @ (M - M.mean(0))
        ev, evec = orc.sym_eig3(cov)
        w, v = np.linalg.eigh(cov)
        assert np.allclose(ev, w, rtol=1e-10, atol=1e-12)
        for c in range(3):
            assert abs(abs(evec[:, c] @ v[:, c]) - 1) < 1e-8
        pts = rng.normal(size=(5, 3)) + [3, 2, 5]
        n, ok = orc.plane_fit5(pts)
        want = np.linalg.lstsq(pts, -np.ones(5), rcond=None)[0]
        assert ok and np.allclose(n, want, rtol=1e-9, atol=1e-11)
