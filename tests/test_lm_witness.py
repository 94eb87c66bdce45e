"""Independent witness for the Levenberg-Marquardt controller (ceres::Solve as configured at LO:819-825 / LM:2079-2087).

`NumpyLM` below is written from SURVEY.md Appendix A.3 and the Ceres documentation only.  It shares no code with
oracle/orc_ceres.cpp or with ll_solve.cuh and deliberately takes different numerical routes:

  * residual Jacobians through the AMBIENT 4-column quaternion Jacobian of Eigen's q * v formula multiplied by the 4 x 3
    EigenQuaternionManifold plus-Jacobian (the chain Ceres itself evaluates) - not the closed form -2 [R cp]x the oracle's
    analytic mode and the CUDA kernel use, and not Jet autodiff;
  * the trust-region step by numpy.linalg.lstsq on the stacked (n + 6) x 6 system - not Householder QR (oracle) and not
    a Cholesky solve of the normal equations (kernel).

The tests replay >= 20 seeded problems through both and compare, iteration by iteration, cost, step norm, relative
decrease, trust-region radius, the valid / successful flags and the termination reason; the problem set is built so that
rejected steps and every exit (max iterations, gradient / parameter / function tolerance) occur, and asserts that they
did.  Also here: witnesses for the Eigen `normalize()` zero-vector branch (LF:210-211) and PCL's "leaf size too small"
fallback (SR:370-374), which the verdict of round 1 listed as unpinned.
"""
import numpy as np

HUBER_A = 0.1


def quat_rotate(q, v):
    """Eigen QuaternionBase::_transformVector: uv = 2 u x v; v + w uv + u x uv (q = x, y, z, w; no normalisation)."""
    u, w = q[:3], q[3]
    uv = 2.0 * np.cross(u, v)
    return v + w * uv + np.cross(u, uv)


def d_rotate_dq(q, v):
    """3 x 4 ambient Jacobian of quat_rotate with respect to (x, y, z, w), differentiated term by term."""
    u, w = q[:3], q[3]
    J = np.zeros((3, 4))
    uxv = np.cross(u, v)
    for k in range(3):
        e = np.zeros(3)
        e[k] = 1.0
        d_uv = 2.0 * np.cross(e, v)
        J[:, k] = w * d_uv + np.cross(e, 2.0 * uxv) + np.cross(u, d_uv)
    J[:, 3] = 2.0 * uxv
    return J


def plus_jacobian(q):
    """EigenQuaternionManifold::PlusJacobian, rows x, y, z, w."""
    x, y, z, w = q
    return np.array([[w, z, -y], [-z, w, x], [y, -x, w], [-x, -y, -z]])


def manifold_plus(x, delta):
    d = delta[:3]
    n = np.linalg.norm(d)
    out = np.array(x, float)
    if n != 0.0:
        dq = np.concatenate([np.sin(n) / n * d, [np.cos(n)]])
        ax, ay, az, aw = dq
        bx, by, bz, bw = x[:4]
        out[:4] = [aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                   aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz]
    out[4:] = x[4:] + delta[3:]
    return out


def block_rows(b):
    return 3 if b[0] == 0 else 1


def evaluate(blocks, x, want_jacobian):
    """Returns (cost, corrected residuals, corrected tangent Jacobian). Block rows as tests/test_oracle_solver.py:
    [type, cp(3), p0(3), p1(3), p2(3), w]: 0 edge (a, b), 1 plane_modify (j, l, m, weight), 2 plane_norm (n, -, -, d)."""
    q, t = x[:4], x[4:]
    res, jac, cost = [], [], 0.0
    PJ = plus_jacobian(q)
    for b in blocks:
        cp = b[1:4]
        lp = quat_rotate(q, cp) + t
        dlp = np.hstack([d_rotate_dq(q, cp) @ PJ, np.eye(3)]) if want_jacobian else None   # 3 x 6 tangent Jacobian of lp
        if b[0] == 0:
            a, c = b[4:7], b[7:10]
            de = np.linalg.norm(a - c)
            r = np.cross(lp - a, lp - c) / de
            if want_jacobian:
                # d/dlp [(lp - a) x (lp - c)] = [ (lp - a) ]x^T ... written out as a finite sum of cross products with unit vectors
                Jr = np.zeros((3, 3))
                for k in range(3):
                    e = np.zeros(3)
                    e[k] = 1.0
                    Jr[:, k] = (np.cross(e, lp - c) + np.cross(lp - a, e)) / de
                J = Jr @ dlp
        elif b[0] == 1:
            j, l, m, w = b[4:7], b[7:10], b[10:13], b[13]
            n = np.cross(j - l, j - m)
            z = float(n @ n)
            if z > 0.0:
                n = n / np.sqrt(z)
            r = np.array([w * float((lp - j) @ n)])
            if want_jacobian:
                J = (w * n)[None, :] @ dlp
        else:
            n, d = b[4:7], b[13]
            r = np.array([float(n @ lp) + d])
            if want_jacobian:
                J = n[None, :] @ dlp
        s = float(r @ r)
        if s > HUBER_A ** 2:          # HuberLoss: rho = 2 a sqrt(s) - a^2, rho' = a / sqrt(s); Corrector with rho'' <= 0: scale by sqrt(rho')
            rho0, rho1 = 2 * HUBER_A * np.sqrt(s) - HUBER_A ** 2, max(np.finfo(float).tiny, HUBER_A / np.sqrt(s))
        else:
            rho0, rho1 = s, 1.0
        cost += 0.5 * rho0
        res.append(np.sqrt(rho1) * r)
        if want_jacobian:
            jac.append(np.sqrt(rho1) * J)
    return cost, np.concatenate(res), (np.vstack(jac) if want_jacobian else None)


class NumpyLM:
    """TrustRegionMinimizer + LevenbergMarquardtStrategy + DENSE_QR with Solver::Options defaults, max_num_iterations = 4."""

    def __init__(self, blocks, x0, max_iters=4):
        self.blocks, self.x, self.max_iters = blocks, np.array(x0, float), max_iters
        self.records = []      # dict per iteration incl. iteration 0
        self.term = None       # 0 max iterations, 1 gradient, 2 parameter, 3 function, 4 radius, 5 failure

    def _linearise(self, first):
        self.cost, self.r, J = evaluate(self.blocks, self.x, True)
        if first:
            self.scale = 1.0 / (1.0 + np.linalg.norm(J, axis=0))
        self.g = J.T @ self.r                      # gradient with the UNSCALED Jacobian (Ceres scales afterwards)
        self.J = J * self.scale
        xp = manifold_plus(self.x, -self.g)
        self.gmax = float(np.abs(self.x - xp).max())

    def run(self):
        radius, decrease, reuse = 1e4, 2.0, False
        self._linearise(True)
        rec = dict(cost=self.cost, gmax=self.gmax, valid=1, successful=1, step_norm=0.0, rel=0.0)
        x_norm = np.linalg.norm(self.x)
        ref_cost = self.cost
        one_ok, invalid, it = False, 0, 0
        while True:
            rec["radius"] = radius
            self.records.append(rec)
            if it >= self.max_iters:
                self.term = 0
                break
            if rec["successful"] and rec["gmax"] <= 1e-10:
                self.term = 1
                break
            if radius <= 1e-32:
                self.term = 4
                break
            prev_gmax = rec["gmax"]
            rec = dict(cost=self.cost, gmax=prev_gmax, valid=0, successful=0, step_norm=0.0, rel=0.0)
            it += 1
            if not reuse:
                diag = np.clip((self.J ** 2).sum(0), 1e-6, 1e32)
            D = np.sqrt(diag / radius)
            A = np.vstack([self.J, np.diag(D)])
            y = np.linalg.lstsq(A, np.concatenate([self.r, np.zeros(6)]), rcond=None)[0]
            reuse = True
            step = -y
            Js = self.J @ step
            model = float(-Js @ (self.r + Js / 2.0))
            if not (np.isfinite(step).all() and model > 0.0):
                invalid += 1
                if invalid >= 5:
                    self.term = 5
                    rec["radius"] = radius
                    self.records.append(rec)
                    break
                radius *= 0.5
                continue
            invalid = 0
            rec["valid"] = 1
            cand = manifold_plus(self.x, step * self.scale)
            cand_cost, _, _ = evaluate(self.blocks, cand, False)
            rec["step_norm"] = float(np.linalg.norm(self.x - cand))
            if one_ok and rec["step_norm"] <= 1e-8 * (x_norm + 1e-8):
                self.term = 2
                break
            rec["cost_change"] = self.cost - cand_cost
            if one_ok and abs(self.cost - cand_cost) <= 1e-6 * self.cost:
                self.term = 3
                break
            rec["rel"] = (ref_cost - cand_cost) / model
            if rec["rel"] > 1e-3:
                self.x = cand
                x_norm = np.linalg.norm(self.x)
                self._linearise(False)
                rec.update(cost=self.cost, gmax=self.gmax, successful=1)
                radius = min(1e16, radius / max(1.0 / 3.0, 1.0 - (2.0 * rec["rel"] - 1.0) ** 3))
                decrease, reuse, ref_cost, one_ok = 2.0, False, cand_cost, True
            else:
                rec.update(cost=cand_cost, successful=0)
                radius /= decrease
                decrease *= 2.0
                reuse = True
        return self


def _rot(q):
    return np.column_stack([quat_rotate(q, e) for e in np.eye(3)])


def make_problem(seed, n=60, pose_err=0.05, noise=0.0, outliers=0, start=None):
    """Blocks consistent with a hidden pose (q*, t*) up to `noise`; the solve starts `pose_err` away from it."""
    rng = np.random.default_rng(seed)
    qs = np.concatenate([rng.normal(size=3) * pose_err, [1.0]])
    qs /= np.linalg.norm(qs)
    ts = rng.normal(size=3) * pose_err * 4
    rows = []
    for i in range(n):
        cp = rng.normal(size=3) * 12
        lp = quat_rotate(qs, cp) + ts + rng.normal(size=3) * noise
        if i < outliers:
            lp = lp + rng.normal(size=3) * 3.0                      # gross mismatches: Huber's outer branch
        k = i % 3
        if k == 0:
            d = rng.normal(size=3)
            rows.append([0, *cp, *(lp + 0.4 * d), *(lp - 0.6 * d), 0, 0, 0, 1.0])
        elif k == 1:
            u, v = rng.normal(size=3), rng.normal(size=3)
            rows.append([1, *cp, *lp, *(lp + u), *(lp + v), [1.0, 5.0][i % 2]])
        else:
            nrm = rng.normal(size=3)
            nrm /= np.linalg.norm(nrm)
            rows.append([2, *cp, *nrm, 0, 0, 0, 0, 0, 0, -float(nrm @ lp)])
    x0 = np.array([0, 0, 0, 1, 0, 0, 0.0]) if start is None else np.array(start, float)
    return np.array(rows, float), x0, np.concatenate([qs, ts])


def make_inconsistent_problem(seed, n, spread, far):
    """Lines and planes that no single pose satisfies (targets drawn independently of the points): a non-convex cost on
    which Gauss-Newton-sized steps overshoot, so the trust region has to reject steps and shrink."""
    rng = np.random.default_rng(seed)
    rows = []
    for i in range(n):
        cp = rng.normal(size=3) * spread
        lp = rng.normal(size=3) * spread * far
        k = i % 3
        if k == 0:
            d = rng.normal(size=3)
            rows.append([0, *cp, *(lp + 0.4 * d), *(lp - 0.6 * d), 0, 0, 0, 1.0])
        elif k == 1:
            u, v = rng.normal(size=3), rng.normal(size=3)
            rows.append([1, *cp, *lp, *(lp + u), *(lp + v), [1.0, 5.0][i % 2]])
        else:
            nrm = rng.normal(size=3)
            nrm /= np.linalg.norm(nrm)
            rows.append([2, *cp, *nrm, 0, 0, 0, 0, 0, 0, -float(nrm @ lp)])
    return np.array(rows, float), np.array([0, 0, 0, 1, 0, 0, 0.0]), None


INCONSISTENT = [dict(seed=s, n=9, spread=0.3, far=3.0) for s in (200, 201, 203, 205, 206)] + \
               [dict(seed=s, n=18, spread=0.05, far=3.0) for s in (201, 202, 207)] + [dict(seed=208, n=9, spread=2.0, far=1.0)]

PROBLEMS = (
    [dict(seed=s, pose_err=0.02, noise=0.0) for s in range(4)] +                 # clean data: converges to ~0 cost -> parameter / gradient exits
    [dict(seed=10 + s, pose_err=0.05, noise=0.03) for s in range(3)] +           # noisy data, Huber active: slow (IRLS-like) convergence
    [dict(seed=50 + s, pose_err=0.02, noise=0.01) for s in range(3)] +           # small noise (quadratic regime): cost plateaus -> function tolerance
    [dict(seed=20 + s, pose_err=0.6, noise=0.02, outliers=12) for s in range(6)] +   # far start + outliers: rejected steps, max iterations
    [dict(seed=30 + s, pose_err=1.2, noise=0.05, outliers=20, n=45) for s in range(4)] +
    [dict(seed=40 + s, pose_err=0.0, noise=0.0) for s in range(2)]               # start at the optimum: gradient tolerance at iteration 0
)


def _oracle_records(orc, blocks, x0):
    x, summ, iters = orc.solve(blocks, x0, max_iters=4, autodiff=True)
    return x, summ, iters


def test_lm_controller_matches_independent_numpy_lm(orc):
    assert len(PROBLEMS) >= 20
    seen_term, saw_rejected, saw_invalid_or_reuse = set(), False, False
    for spec in list(PROBLEMS) + INCONSISTENT:
        blocks, x0, _ = make_inconsistent_problem(**spec) if "far" in spec else make_problem(**spec)
        w = NumpyLM(blocks, x0).run()
        x, summ, iters = _oracle_records(orc, blocks, x0)
        assert int(summ[5]) == w.term, (spec, summ[5], w.term)
        assert len(iters) == len(w.records), (spec, len(iters), len(w.records))
        for k, (o, r) in enumerate(zip(iters, w.records)):
            # oracle row: cost, cost_change, gradient_max_norm, step_norm, relative_decrease, radius, valid, successful
            assert int(o[6]) == r["valid"] and int(o[7]) == r["successful"], (spec, k)
            assert np.isclose(o[0], r["cost"], rtol=1e-8, atol=1e-18), (spec, k, o[0], r["cost"])
            assert np.isclose(o[5], r["radius"], rtol=1e-6), (spec, k, o[5], r["radius"])
            assert np.isclose(o[3], r["step_norm"], rtol=1e-6, atol=1e-14), (spec, k)
            if r["valid"] and k > 0:
                assert np.isclose(o[4], r["rel"], rtol=1e-5, atol=1e-9), (spec, k, o[4], r["rel"])
            if r["successful"]:
                assert np.isclose(o[2], r["gmax"], rtol=1e-5, atol=1e-11), (spec, k)   # near the optimum the gradient is cancellation noise
            saw_rejected |= (k > 0 and r["valid"] == 1 and r["successful"] == 0)
        assert np.allclose(x, w.x, rtol=1e-9, atol=1e-11), spec
        assert np.isclose(summ[1], w.cost, rtol=1e-8, atol=1e-18)
        seen_term.add(w.term)
    assert saw_rejected, "no rejected step in the problem set"
    assert {0, 1, 2, 3} <= seen_term, seen_term           # max iterations, gradient, parameter and function tolerance all exercised


def test_witness_jacobian_chain_equals_oracle_autodiff(orc):
    """Ambient Jacobian x plus-Jacobian (this file) == the oracle's Jet autodiff == its closed form, on Huber's both branches."""
    blocks, _, _ = make_problem(3, n=30, pose_err=0.3, noise=0.05, outliers=6)
    q = np.array([0.07, -0.03, 0.05, 1.0])
    q /= np.linalg.norm(q)
    x = np.concatenate([q, [0.3, -0.2, 0.1]])
    cost, r, J = evaluate(blocks, x, True)
    oc, orr, og, oJ = orc.evaluate(blocks, x, autodiff=True)
    _, _, _, oJ2 = orc.evaluate(blocks, x, autodiff=False)
    assert np.isclose(cost, oc, rtol=1e-12)
    assert np.allclose(r, orr, rtol=1e-11, atol=1e-13)
    assert np.allclose(J, oJ, rtol=1e-9, atol=1e-11) and np.allclose(J, oJ2, rtol=1e-9, atol=1e-11)
    assert np.allclose(J.T @ r, og, rtol=1e-9, atol=1e-11)


def test_eigen_normalize_zero_vector_branch(orc):
    """LF:210-211: ljm_norm = (j - l).cross(j - m); normalize().  Eigen's normalize() leaves a zero vector untouched
    (z = squaredNorm(); if (z > 0) *this /= sqrt(z)), so a degenerate triple (collinear j, l, m) yields a zero normal:
    residual 0, Jacobian row 0, no NaN - in this file's NumPy replica and in the oracle."""
    cp = np.array([1.0, 2.0, 3.0])
    j = np.array([4.0, 5.0, 6.0])
    rows = np.array([[1, *cp, *j, *(j + [1, 1, 1]), *(j + [2, 2, 2]), 5.0],          # collinear: cross product exactly 0
                     [1, *cp, *j, *(j + [1, 0, 0]), *(j + [0, 1, 0]), 1.0]], float)  # a regular one next to it
    x = np.array([0.01, 0.02, -0.01, 1.0, 0.1, 0.2, 0.3])
    x[:4] /= np.linalg.norm(x[:4])
    cost, r, J = evaluate(rows, x, True)
    oc, orr, og, oJ = orc.evaluate(rows, x, autodiff=True)
    assert r[0] == 0.0 and not J[0].any() and np.isfinite(J).all()
    assert orr[0] == 0.0 and not oJ[0].any() and np.isfinite(oJ).all()
    assert np.allclose(r, orr, rtol=1e-12) and np.allclose(J, oJ, rtol=1e-10, atol=1e-12) and np.isclose(cost, oc)


def _np_voxel_grid(cloud, leaf):
    """pcl::VoxelGrid::applyFilter from SURVEY A.1 in plain NumPy / Python (stable within-voxel order)."""
    inv = np.float32(1.0) / np.float32(leaf)
    xyz = cloud[:, :3].astype(np.float32)
    mn, mx = xyz.min(0), xyz.max(0)
    d = [int(np.int64(np.float32(mx[a] - mn[a]) * inv)) + 1 for a in range(3)]
    if d[0] * d[1] * d[2] > np.iinfo(np.int32).max:
        return cloud.copy(), True
    min_b = np.floor(mn * inv).astype(np.int64)
    max_b = np.floor(mx * inv).astype(np.int64)
    div = max_b - min_b + 1
    ijk = (np.floor(xyz * inv) - min_b.astype(np.float32)).astype(np.int64)
    idx = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    out = []
    for v in np.unique(idx):
        acc = np.zeros(4, np.float32)
        members = np.nonzero(idx == v)[0]
        for m in members:
            acc = (acc + cloud[m]).astype(np.float32)
        out.append(acc / np.float32(len(members)))
    return np.array(out, np.float32), False


def test_pcl_leaf_too_small_fallback_and_its_threshold(orc):
    """dx * dy * dz > INT32_MAX -> "Leaf size is too small", output = input (order kept); one step below the threshold the
    filter runs normally.  The NumPy restatement above decides the branch independently of the oracle."""
    rng = np.random.default_rng(11)
    cloud = np.zeros((400, 4), np.float32)
    cloud[:, :3] = rng.uniform(-1, 1, size=(400, 3)).astype(np.float32)
    cloud[:, 3] = rng.uniform(0, 63, size=400).astype(np.float32)
    cloud[0, :3] = [-130.0, -130.0, -130.0]
    cloud[1, :3] = [130.0, 130.0, 130.0]           # extent 260 m: (260 / 0.2 + 1)^3 = 2.2e9 > INT32_MAX
    want, small = _np_voxel_grid(cloud, 0.2)
    assert small
    got = orc.voxel_grid(cloud, 0.2, stable=True)
    assert np.array_equal(got, cloud) and np.array_equal(want, cloud)
    cloud[0, :3] = [-125.0, -125.0, -125.0]
    cloud[1, :3] = [125.0, 125.0, 125.0]           # (250 / 0.2 + 1)^3 = 1.96e9 < INT32_MAX: filtered
    want, small = _np_voxel_grid(cloud, 0.2)
    assert not small
    got = orc.voxel_grid(cloud, 0.2, stable=True)
    assert got.shape == want.shape and len(got) < len(cloud)
    assert np.array_equal(got, want)
