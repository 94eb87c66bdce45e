// ORACLE (test infrastructure) — the reference's cost functors (lidarFactor.hpp) and ceres::Solve as the
// reference configures it, restated without Ceres/Eigen.
//
// Live functors: LidarEdgeFactor LF:9-52 (LO:615, LM:1918), LidarPlaneFactor_modify LF:203-251
// (LO:783, LO:804), LidarPlaneNormFactor LF:253-285 (LM:2033).  Solver set-up: HuberLoss(0.1),
// EigenQuaternionManifold on the 4-block, DENSE_QR, max_num_iterations = 4 (LO:475-482, 819-825;
// LM:1865-1872, 2079-2087), every other option at its Ceres 2.x default.
//
// Ceres is an un-vendored dependency (CMakeLists.txt:24 pins 2.3); restated from the published
// sources of Ceres 2.x: trust_region_minimizer.cc, levenberg_marquardt_strategy.cc,
// trust_region_step_evaluator.cc, residual_block.cc, corrector.cc, loss_function.cc (HuberLoss),
// manifold.cc (EigenQuaternionManifold), dense_qr_solver.cc, jet.h — SURVEY.md Appendix A.3.
#include "orc_api.h"
#include "orc_jet.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>

namespace orc {

ResidualBlock make_edge(const double cp[3], const double a[3], const double b[3], double s)
{
    ResidualBlock r{};
    r.type = EDGE;
    for (int i = 0; i < 3; ++i) { r.cp[i] = cp[i]; r.a[i] = a[i]; r.b[i] = b[i]; }
    r.s = s;
    r.w = 1.0;
    return r;
}

ResidualBlock make_plane_modify(const double cp[3], const double j[3], const double l[3], const double m[3], double s, double weight)
{
    ResidualBlock r{};
    r.type = PLANE_MODIFY;
    // LF:210-211  ljm_norm = (j - l).cross(j - m); ljm_norm.normalize();
    const V3<double> jl{j[0] - l[0], j[1] - l[1], j[2] - l[2]}, jm{j[0] - m[0], j[1] - m[1], j[2] - m[2]};
    V3<double> n = cross(jl, jm);
    const double z = dot(n, n);  // Eigen normalize(): z = squaredNorm(); if (z > 0) *this /= sqrt(z)
    if (z > 0.0) { const double nn = std::sqrt(z); n = {n.x / nn, n.y / nn, n.z / nn}; }
    for (int i = 0; i < 3; ++i) { r.cp[i] = cp[i]; r.a[i] = j[i]; }
    r.b[0] = n.x; r.b[1] = n.y; r.b[2] = n.z;
    r.s = s;
    r.w = weight;
    return r;
}

ResidualBlock make_plane_norm(const double cp[3], const double n[3], double d)
{
    ResidualBlock r{};
    r.type = PLANE_NORM;
    for (int i = 0; i < 3; ++i) { r.cp[i] = cp[i]; r.a[i] = n[i]; r.b[i] = 0.0; }
    r.s = 1.0;
    r.w = d;
    return r;
}

namespace {

// The functors' operator() for T = double or Jet<7>; q = (x,y,z,w) ambient, t = translation.
template <typename T>
int functor(const ResidualBlock& blk, const T* q, const T* t, T* residual)
{
    const V3<T> cp{T(blk.cp[0]), T(blk.cp[1]), T(blk.cp[2])};
    if (blk.type == PLANE_NORM) {  // LF:259-271
        const Quat<T> q_w_curr{q[0], q[1], q[2], q[3]};
        const V3<T> t_w_curr{t[0], t[1], t[2]};
        const V3<T> point_w = rotate(q_w_curr, cp) + t_w_curr;
        const V3<T> nrm{T(blk.a[0]), T(blk.a[1]), T(blk.a[2])};
        residual[0] = dot(nrm, point_w) + T(blk.w);
        return 1;
    }
    // LF:23-31 / LF:227-235: q_last_curr = identity.slerp(T(s), q); t_last_curr = T(s) * t
    Quat<T> q_last_curr{q[0], q[1], q[2], q[3]};
    q_last_curr = identity_slerp(T(blk.s), q_last_curr);
    const V3<T> t_last_curr{T(blk.s) * t[0], T(blk.s) * t[1], T(blk.s) * t[2]};
    const V3<T> lp = rotate(q_last_curr, cp) + t_last_curr;
    if (blk.type == EDGE) {  // LF:33-38
        const V3<T> lpa{T(blk.a[0]), T(blk.a[1]), T(blk.a[2])}, lpb{T(blk.b[0]), T(blk.b[1]), T(blk.b[2])};
        const V3<T> nu = cross(lp - lpa, lp - lpb);
        const V3<T> de = lpa - lpb;
        residual[0] = nu.x / norm(de);
        residual[1] = nu.y / norm(de);
        residual[2] = nu.z / norm(de);
        return 3;
    }
    // PLANE_MODIFY, LF:237
    const V3<T> lpj{T(blk.a[0]), T(blk.a[1]), T(blk.a[2])}, ljm{T(blk.b[0]), T(blk.b[1]), T(blk.b[2])};
    residual[0] = dot(lp - lpj, ljm) * T(blk.w);
    return 1;
}

// Closed-form tangent Jacobian (SURVEY.md §8a "derived simplification"): used only by the tests as an
// independent check of the autodiff path (use_autodiff = false).
int analytic(const ResidualBlock& blk, const double* x, double* res, double* J /* rows x 6 */)
{
    const Quat<double> q{x[0], x[1], x[2], x[3]};
    const V3<double> cp{blk.cp[0], blk.cp[1], blk.cp[2]};
    const V3<double> Rp = rotate(q, cp);
    const V3<double> lp{Rp.x + x[4], Rp.y + x[5], Rp.z + x[6]};
    // d lp / d delta = -2 [R cp]x ; d lp / d t = I
    const double M[3][6] = {{0, 2 * Rp.z, -2 * Rp.y, 1, 0, 0}, {-2 * Rp.z, 0, 2 * Rp.x, 0, 1, 0}, {2 * Rp.y, -2 * Rp.x, 0, 0, 0, 1}};
    if (blk.type == EDGE) {
        const V3<double> a{blk.a[0], blk.a[1], blk.a[2]}, b{blk.b[0], blk.b[1], blk.b[2]};
        const V3<double> nu = cross(lp - a, lp - b), de = a - b;
        const double dn = norm(de);
        res[0] = nu.x / dn; res[1] = nu.y / dn; res[2] = nu.z / dn;
        const V3<double> e = b - a;  // d nu / d lp = [b - a]x
        const double D[3][3] = {{0, -e.z / dn, e.y / dn}, {e.z / dn, 0, -e.x / dn}, {-e.y / dn, e.x / dn, 0}};
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 6; ++c) J[r * 6 + c] = D[r][0] * M[0][c] + D[r][1] * M[1][c] + D[r][2] * M[2][c];
        return 3;
    }
    double n[3], w = 1.0;
    if (blk.type == PLANE_MODIFY) {
        n[0] = blk.b[0]; n[1] = blk.b[1]; n[2] = blk.b[2];
        w = blk.w;
        res[0] = ((lp.x - blk.a[0]) * n[0] + (lp.y - blk.a[1]) * n[1] + (lp.z - blk.a[2]) * n[2]) * w;
    } else {
        n[0] = blk.a[0]; n[1] = blk.a[1]; n[2] = blk.a[2];
        res[0] = n[0] * lp.x + n[1] * lp.y + n[2] * lp.z + blk.w;
    }
    for (int c = 0; c < 6; ++c) J[c] = w * (n[0] * M[0][c] + n[1] * M[1][c] + n[2] * M[2][c]);
    return 1;
}

}  // namespace

// EigenQuaternionManifold::Plus on the first block, Euclidean on the second (manifold.cc).
void manifold_plus(const double x[7], const double delta[6], double out[7])
{
    const double norm_delta = std::sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
    if (norm_delta == 0.0) {
        for (int i = 0; i < 4; ++i) out[i] = x[i];
    } else {
        const double sin_delta_by_delta = std::sin(norm_delta) / norm_delta;
        const Quat<double> dq{sin_delta_by_delta * delta[0], sin_delta_by_delta * delta[1], sin_delta_by_delta * delta[2],
                              std::cos(norm_delta)};
        const Quat<double> r = qmul(dq, Quat<double>{x[0], x[1], x[2], x[3]});
        out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
    }
    for (int i = 0; i < 3; ++i) out[4 + i] = x[4 + i] + delta[3 + i];
}

// ResidualBlock::Evaluate + ProgramEvaluator: cost, corrected residuals, tangent Jacobian, gradient.
int evaluate(const std::vector<ResidualBlock>& blocks, const double x[7], double* cost, std::vector<double>* residuals,
             double* gradient6, std::vector<double>* jacobian, bool use_autodiff)
{
    const bool want_j = jacobian != nullptr || gradient6 != nullptr;
    int rows = 0;
    for (const ResidualBlock& b : blocks) rows += b.type == EDGE ? 3 : 1;
    if (residuals) residuals->assign(rows, 0.0);
    if (jacobian) jacobian->assign((size_t)rows * 6, 0.0);
    if (gradient6) for (int i = 0; i < 6; ++i) gradient6[i] = 0.0;
    // EigenQuaternionManifold::PlusJacobian, rows (x,y,z,w)
    const double PJ[4][3] = {{x[3], x[2], -x[1]}, {-x[2], x[3], x[0]}, {x[1], -x[0], x[3]}, {-x[0], -x[1], -x[2]}};
    double total = 0.0;
    int row = 0;
    for (const ResidualBlock& blk : blocks) {
        double r[3];
        double J[18];
        int nr;
        if (!want_j) {
            nr = functor<double>(blk, x, x + 4, r);
        } else if (use_autodiff) {
            Jet<7> jq[4], jt[3], jr[3];
            for (int i = 0; i < 4; ++i) jq[i] = Jet<7>(x[i], i);
            for (int i = 0; i < 3; ++i) jt[i] = Jet<7>(x[4 + i], 4 + i);
            nr = functor<Jet<7>>(blk, jq, jt, jr);
            for (int k = 0; k < nr; ++k) {
                r[k] = jr[k].a;
                for (int c = 0; c < 3; ++c)  // global_jacobian(4) * plus_jacobian(4x3)
                    J[k * 6 + c] = jr[k].v[0] * PJ[0][c] + jr[k].v[1] * PJ[1][c] + jr[k].v[2] * PJ[2][c] + jr[k].v[3] * PJ[3][c];
                for (int c = 0; c < 3; ++c) J[k * 6 + 3 + c] = jr[k].v[4 + c];
            }
        } else {
            nr = analytic(blk, x, r, J);
        }
        double squared_norm = 0.0;
        for (int k = 0; k < nr; ++k) squared_norm += r[k] * r[k];
        // HuberLoss(a = 0.1)::Evaluate
        const double a_ = 0.1, b_ = a_ * a_;
        double rho0, rho1;
        if (squared_norm > b_) {
            const double rr = std::sqrt(squared_norm);
            rho0 = 2.0 * a_ * rr - b_;
            rho1 = std::max(std::numeric_limits<double>::min(), a_ / rr);
        } else {
            rho0 = squared_norm;
            rho1 = 1.0;
        }
        total += 0.5 * rho0;
        // Corrector: rho'' <= 0 always for Huber -> residual_scaling = sqrt(rho'), alpha = 0
        const double sqrt_rho1 = std::sqrt(rho1);
        if (want_j) {
            for (int k = 0; k < nr * 6; ++k) J[k] *= sqrt_rho1;
        }
        for (int k = 0; k < nr; ++k) r[k] *= sqrt_rho1;
        if (residuals) for (int k = 0; k < nr; ++k) (*residuals)[row + k] = r[k];
        if (jacobian) std::memcpy(jacobian->data() + (size_t)row * 6, J, sizeof(double) * nr * 6);
        if (gradient6)
            for (int k = 0; k < nr; ++k)
                for (int c = 0; c < 6; ++c) gradient6[c] += J[k * 6 + c] * r[k];
        row += nr;
    }
    *cost = total;
    return rows;
}

namespace {

// Householder QR least squares: min || A y - b ||, A is m x 6 row-major (destroyed), b length m (destroyed).
bool qr_solve6(std::vector<double>& A, std::vector<double>& b, int m, double y[6])
{
    const int n = 6;
    for (int k = 0; k < n; ++k) {
        double nrm = 0.0;
        for (int i = k; i < m; ++i) nrm += A[(size_t)i * n + k] * A[(size_t)i * n + k];
        nrm = std::sqrt(nrm);
        if (nrm == 0.0) return false;
        const double alpha = A[(size_t)k * n + k] > 0 ? -nrm : nrm;
        const double v0 = A[(size_t)k * n + k] - alpha;
        // v = (v0, A[k+1..m, k]); H = I - 2 v v^T / (v^T v)
        double vtv = v0 * v0;
        for (int i = k + 1; i < m; ++i) vtv += A[(size_t)i * n + k] * A[(size_t)i * n + k];
        if (vtv == 0.0) continue;
        for (int j = k + 1; j < n; ++j) {
            double s = v0 * A[(size_t)k * n + j];
            for (int i = k + 1; i < m; ++i) s += A[(size_t)i * n + k] * A[(size_t)i * n + j];
            s = 2.0 * s / vtv;
            A[(size_t)k * n + j] -= s * v0;
            for (int i = k + 1; i < m; ++i) A[(size_t)i * n + j] -= s * A[(size_t)i * n + k];
        }
        {
            double s = v0 * b[k];
            for (int i = k + 1; i < m; ++i) s += A[(size_t)i * n + k] * b[i];
            s = 2.0 * s / vtv;
            b[k] -= s * v0;
            for (int i = k + 1; i < m; ++i) b[i] -= s * A[(size_t)i * n + k];
        }
        A[(size_t)k * n + k] = alpha;
    }
    for (int k = n - 1; k >= 0; --k) {
        double s = b[k];
        for (int j = k + 1; j < n; ++j) s -= A[(size_t)k * n + j] * y[j];
        y[k] = s / A[(size_t)k * n + k];
        if (!std::isfinite(y[k])) return false;
    }
    return true;
}

}  // namespace

void solve(const std::vector<ResidualBlock>& blocks, double q[4], double t[3], SolveSummary* summary_out,
           int max_num_iterations, bool use_autodiff)
{
    SolveSummary S;
    double x[7] = {q[0], q[1], q[2], q[3], t[0], t[1], t[2]};
    // Ceres: a problem with no residual blocks has nothing to minimise; parameters are left untouched.
    if (blocks.empty()) { if (summary_out) *summary_out = S; return; }

    // Solver::Options defaults
    const double function_tolerance = 1e-6, gradient_tolerance = 1e-10, parameter_tolerance = 1e-8;
    const double min_relative_decrease = 1e-3, min_lm_diagonal = 1e-6, max_lm_diagonal = 1e32;
    const double max_trust_region_radius = 1e16, min_trust_region_radius = 1e-32;
    const int max_num_consecutive_invalid_steps = 5;
    double radius = 1e4, decrease_factor = 2.0;
    bool reuse_diagonal = false;

    std::vector<double> residuals, jacobian;
    double gradient[6], scale[6], diagonal[6];
    double x_cost = 0.0, x_norm;
    auto norm7 = [](const double* v) { double s = 0; for (int i = 0; i < 7; ++i) s += v[i] * v[i]; return std::sqrt(s); };

    int rows = 0;
    IterRecord it{};
    // EvaluateGradientAndJacobian(): evaluate, (iteration 0: compute Jacobi scaling), scale columns, projected gradient
    auto eval_gj = [&](bool first) {
        rows = evaluate(blocks, x, &x_cost, &residuals, gradient, &jacobian, use_autodiff);
        S.num_jacobian_evals++;
        if (first) {
            for (int c = 0; c < 6; ++c) {
                double s = 0.0;
                for (int r = 0; r < rows; ++r) s += jacobian[(size_t)r * 6 + c] * jacobian[(size_t)r * 6 + c];
                scale[c] = 1.0 / (1.0 + std::sqrt(s));
            }
        }
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < 6; ++c) jacobian[(size_t)r * 6 + c] *= scale[c];
        double neg_g[6], xp[7];
        for (int c = 0; c < 6; ++c) neg_g[c] = -gradient[c];
        manifold_plus(x, neg_g, xp);
        double mx = 0.0;
        for (int i = 0; i < 7; ++i) mx = std::max(mx, std::fabs(x[i] - xp[i]));
        it.gradient_max_norm = mx;
        it.cost = x_cost;
    };

    // IterationZero()
    x_norm = norm7(x);
    eval_gj(true);
    S.initial_cost = x_cost;
    it.valid = 1;
    it.successful = 1;
    double se_current_cost = x_cost;  // TrustRegionStepEvaluator (monotonic): reference == current
    int num_consecutive_invalid_steps = 0;
    bool atleast_one_successful_step = false;
    int iteration = 0;
    double prev_gmax = it.gradient_max_norm;
    S.termination = 0;

    for (;;) {
        // FinalizeIterationAndCheckIfMinimizerCanContinue()
        it.radius = radius;
        S.iterations.push_back(it);
        if (iteration >= max_num_iterations) { S.termination = 0; break; }
        if (it.successful && it.gradient_max_norm <= gradient_tolerance) { S.termination = 1; break; }
        if (radius <= min_trust_region_radius) { S.termination = 4; break; }

        prev_gmax = it.gradient_max_norm;
        it = IterRecord{};
        ++iteration;

        // ComputeTrustRegionStep(): LevenbergMarquardtStrategy::ComputeStep + DenseQRSolver
        if (!reuse_diagonal) {
            for (int c = 0; c < 6; ++c) {
                double s = 0.0;
                for (int r = 0; r < rows; ++r) s += jacobian[(size_t)r * 6 + c] * jacobian[(size_t)r * 6 + c];
                diagonal[c] = std::min(std::max(s, min_lm_diagonal), max_lm_diagonal);
            }
        }
        double lm_diagonal[6];
        for (int c = 0; c < 6; ++c) lm_diagonal[c] = std::sqrt(diagonal[c] / radius);
        std::vector<double> A((size_t)(rows + 6) * 6, 0.0), b(rows + 6, 0.0);
        std::memcpy(A.data(), jacobian.data(), sizeof(double) * (size_t)rows * 6);
        for (int c = 0; c < 6; ++c) A[(size_t)(rows + c) * 6 + c] = lm_diagonal[c];
        std::memcpy(b.data(), residuals.data(), sizeof(double) * rows);
        double step[6];
        const bool ok = qr_solve6(A, b, rows + 6, step);  // solves J y = r; step = -y
        reuse_diagonal = true;
        bool step_is_valid = false;
        double model_cost_change = 0.0;
        double delta[6];
        if (ok) {
            for (int c = 0; c < 6; ++c) step[c] = -step[c];
            // model_cost_change = -(J step)' (r + J step / 2)
            for (int r = 0; r < rows; ++r) {
                double m = 0.0;
                for (int c = 0; c < 6; ++c) m += jacobian[(size_t)r * 6 + c] * step[c];
                model_cost_change += -m * (residuals[r] + m / 2.0);
            }
            step_is_valid = model_cost_change > 0.0;
        }
        it.valid = step_is_valid;
        if (!step_is_valid) {  // HandleInvalidStep()
            if (++num_consecutive_invalid_steps >= max_num_consecutive_invalid_steps) {
                S.termination = 5;
                it.cost = x_cost;
                it.radius = radius;
                S.iterations.push_back(it);
                break;
            }
            radius *= 0.5;  // StepIsInvalid()
            reuse_diagonal = true;
            it.cost = x_cost;
            it.gradient_max_norm = prev_gmax;
            continue;
        }
        num_consecutive_invalid_steps = 0;
        for (int c = 0; c < 6; ++c) delta[c] = step[c] * scale[c];

        // ComputeCandidatePointAndEvaluateCost()
        double cand[7], cand_cost;
        manifold_plus(x, delta, cand);
        evaluate(blocks, cand, &cand_cost, nullptr, nullptr, nullptr, use_autodiff);
        S.num_cost_evals++;

        // ParameterToleranceReached()
        double sn = 0.0;
        for (int i = 0; i < 7; ++i) sn += (x[i] - cand[i]) * (x[i] - cand[i]);
        it.step_norm = std::sqrt(sn);
        if (atleast_one_successful_step && it.step_norm <= parameter_tolerance * (x_norm + parameter_tolerance)) {
            S.termination = 2;
            break;
        }
        // FunctionToleranceReached()
        it.cost_change = x_cost - cand_cost;
        if (atleast_one_successful_step && std::fabs(it.cost_change) <= function_tolerance * x_cost) {
            S.termination = 3;
            break;
        }
        // IsStepSuccessful(): StepQuality (monotonic steps)
        it.relative_decrease = (se_current_cost - cand_cost) / model_cost_change;
        if (it.relative_decrease > min_relative_decrease) {
            // HandleSuccessfulStep()
            for (int i = 0; i < 7; ++i) x[i] = cand[i];
            x_norm = norm7(x);
            eval_gj(false);
            it.successful = 1;
            // StepAccepted(step_quality)
            radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * it.relative_decrease - 1.0, 3));
            radius = std::min(max_trust_region_radius, radius);
            decrease_factor = 2.0;
            reuse_diagonal = false;
            se_current_cost = cand_cost;
            atleast_one_successful_step = true;
        } else {
            it.successful = 0;
            it.cost = cand_cost;
            it.gradient_max_norm = prev_gmax;
            radius = radius / decrease_factor;  // StepRejected()
            decrease_factor *= 2.0;
            reuse_diagonal = true;
        }
    }
    // the minimizer writes x to the user's parameters on every successful, cost-decreasing step;
    // with monotonic steps that is the current x.
    for (int i = 0; i < 4; ++i) q[i] = x[i];
    for (int i = 0; i < 3; ++i) t[i] = x[4 + i];
    S.final_cost = x_cost;
    S.num_iterations = (int)S.iterations.size();
    if (summary_out) *summary_out = S;
}

}  // namespace orc
