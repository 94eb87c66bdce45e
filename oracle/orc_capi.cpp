// ORACLE (test infrastructure) — C ABI over the CPU restatement, for ctypes (tests/, smoke(), and
// bench.py's cpu_baseline / --impl reference legs only).  Never linked into the product library.
#include "orc_api.h"

#include <chrono>
#include <cstring>

using namespace orc;

extern "C" {

struct orc_config {
    int scan_line;
    float minimum_range, lower_bound, up_bound, line_res, plane_res;
    int skip_frame, voxel_stable, graph_from_frame;
    int map_graph_vote, distortion, vote_mode;
};

static Config to_cfg(const orc_config* c)
{
    Config k;
    if (!c) return k;
    k.scan_line = c->scan_line; k.minimum_range = c->minimum_range; k.lower_bound = c->lower_bound; k.up_bound = c->up_bound;
    k.line_res = c->line_res; k.plane_res = c->plane_res; k.skip_frame = c->skip_frame; k.voxel_stable = c->voxel_stable;
    k.graph_from_frame = c->graph_from_frame;
    k.map_graph_vote = c->map_graph_vote; k.distortion = c->distortion; k.vote_mode = c->vote_mode;
    return k;
}

static std::vector<P4> to_cloud(const float* p, int n) { std::vector<P4> v(n > 0 ? n : 0); if (n > 0) std::memcpy(v.data(), p, sizeof(P4) * n); return v; }

// ---- features -----------------------------------------------------------------------------------------
void* orc_features_run(const float* pts, int n, int stride_floats, const orc_config* cfg, int* rc)
{
    Features* f = new Features();
    const int r = extract_features(pts, n, stride_floats, to_cfg(cfg), *f);
    if (rc) *rc = r;
    return f;
}
// sizes: n_full, n_sharp, n_less_sharp, n_flat, n_less_flat, n_rings, sort_ties
void orc_features_sizes(void* h, long long out[7])
{
    Features* f = (Features*)h;
    out[0] = (long long)f->full.size(); out[1] = (long long)f->sharp_idx.size(); out[2] = (long long)f->less_sharp_idx.size();
    out[3] = (long long)f->flat_idx.size(); out[4] = (long long)f->less_flat.size();
    out[5] = (long long)f->less_flat_ring_count.size(); out[6] = f->sort_ties;
}
void orc_features_copy(void* h, float* full, int* ring_begin, float* curvature, int* label, int* sharp_idx, int* less_sharp_idx,
                       int* flat_idx, float* less_flat, int* less_flat_ring_count)
{
    Features* f = (Features*)h;
    if (full) std::memcpy(full, f->full.data(), sizeof(P4) * f->full.size());
    if (ring_begin) std::memcpy(ring_begin, f->ring_begin.data(), sizeof(int) * f->ring_begin.size());
    if (curvature) std::memcpy(curvature, f->curvature.data(), sizeof(float) * f->curvature.size());
    if (label) std::memcpy(label, f->label.data(), sizeof(int) * f->label.size());
    if (sharp_idx) std::memcpy(sharp_idx, f->sharp_idx.data(), sizeof(int) * f->sharp_idx.size());
    if (less_sharp_idx) std::memcpy(less_sharp_idx, f->less_sharp_idx.data(), sizeof(int) * f->less_sharp_idx.size());
    if (flat_idx) std::memcpy(flat_idx, f->flat_idx.data(), sizeof(int) * f->flat_idx.size());
    if (less_flat) std::memcpy(less_flat, f->less_flat.data(), sizeof(P4) * f->less_flat.size());
    if (less_flat_ring_count) std::memcpy(less_flat_ring_count, f->less_flat_ring_count.data(), sizeof(int) * f->less_flat_ring_count.size());
}
void orc_features_free(void* h) { delete (Features*)h; }

// ---- voxel grid / k-NN / small dense -----------------------------------------------------------------
int orc_voxel_grid(const float* in, int n, float leaf, int stable, float* out, int cap)
{
    std::vector<P4> o;
    voxel_grid(to_cloud(in, n), leaf, stable != 0, o);
    if ((int)o.size() > cap) return -(int)o.size();
    std::memcpy(out, o.data(), sizeof(P4) * o.size());
    return (int)o.size();
}
// exact k-NN of nq queries (xyz float3 packed) in a cloud (float4); idx/d2 are nq*k, -1 / inf padded
void orc_knn(const float* cloud, int n, const float* queries, int nq, int k, int* idx, float* d2)
{
    KdTree t;
    t.build(to_cloud(cloud, n));
    for (int i = 0; i < nq; ++i) {
        for (int j = 0; j < k; ++j) { idx[(size_t)i * k + j] = -1; d2[(size_t)i * k + j] = 3.0e38f; }
        t.knn(queries + 3 * (size_t)i, k, idx + (size_t)i * k, d2 + (size_t)i * k);
    }
}
void orc_sym_eig3(const double A[9], double evals[3], double evecs[9]) { sym_eig3(A, evals, evecs); }
int orc_plane_fit5(const double pts[15], double n[3]) { return plane_fit5(pts, n) ? 1 : 0; }

// ---- graph vote ---------------------------------------------------------------------------------------
// src/tgt: n float4 each. Outputs: votes[n]; selected index / weight arrays (cap n); returns #selected.
int orc_graph_vote(const float* src, const float* tgt, int n, int corner_case, float* votes, int* sel_idx, float* sel_w)
{
    std::vector<CorreMatch> c(n);
    for (int i = 0; i < n; ++i) {
        c[i].index = i;
        std::memcpy(&c[i].src, src + 4 * (size_t)i, sizeof(P4));
        std::memcpy(&c[i].tgt, tgt + 4 * (size_t)i, sizeof(P4));
        c[i].score = 0; c[i].s = 1;
    }
    std::vector<VertexVote> sel;
    std::vector<float> v;
    graph_vote_simple(c, corner_case != 0, sel, &v);
    if (votes) std::memcpy(votes, v.data(), sizeof(float) * n);
    for (size_t i = 0; i < sel.size(); ++i) { if (sel_idx) sel_idx[i] = sel[i].index; if (sel_w) sel_w[i] = sel[i].score; }
    return (int)sel.size();
}

// ---- residual blocks / solver -------------------------------------------------------------------------
// blocks: n x 14 doubles = {type, cp[3], p0[3], p1[3], p2[3], weight}:
//   type 0 EDGE: p0 = a, p1 = b | type 1 PLANE_MODIFY: p0 = j, p1 = l, p2 = m, weight | type 2 PLANE_NORM: p0 = n, weight = d
static std::vector<ResidualBlock> to_blocks(const double* b, int n)
{
    std::vector<ResidualBlock> v;
    v.reserve(n);
    for (int i = 0; i < n; ++i) {
        const double* r = b + 14 * (size_t)i;
        const int type = (int)r[0];
        if (type == EDGE) v.push_back(make_edge(r + 1, r + 4, r + 7, 1.0));
        else if (type == PLANE_MODIFY) v.push_back(make_plane_modify(r + 1, r + 4, r + 7, r + 10, 1.0, r[13]));
        else v.push_back(make_plane_norm(r + 1, r + 4, r[13]));
    }
    return v;
}
// returns #rows; residuals (rows), gradient (6), jacobian (rows*6) optional
int orc_evaluate(const double* blocks, int n, const double x[7], int use_autodiff, double* cost, double* residuals, double* gradient,
                 double* jacobian)
{
    std::vector<double> r, J;
    const int rows = evaluate(to_blocks(blocks, n), x, cost, residuals ? &r : nullptr, gradient, jacobian ? &J : nullptr, use_autodiff != 0);
    if (residuals) std::memcpy(residuals, r.data(), sizeof(double) * r.size());
    if (jacobian) std::memcpy(jacobian, J.data(), sizeof(double) * J.size());
    return rows;
}
void orc_manifold_plus(const double x[7], const double delta[6], double out[7]) { manifold_plus(x, delta, out); }
// x in/out (q xyzw, t). summary: {initial_cost, final_cost, num_iterations, jac_evals, cost_evals, termination}
// iters: up to 8 records x 8 doubles {cost, cost_change, gmax, step_norm, rel_decrease, radius, valid, successful}
void orc_solve(const double* blocks, int n, double x[7], int max_iters, int use_autodiff, double summary[6], double* iters)
{
    SolveSummary s;
    solve(to_blocks(blocks, n), x, x + 4, &s, max_iters, use_autodiff != 0);
    if (summary) { summary[0] = s.initial_cost; summary[1] = s.final_cost; summary[2] = s.num_iterations; summary[3] = s.num_jacobian_evals; summary[4] = s.num_cost_evals; summary[5] = s.termination; }
    if (iters)
        for (size_t i = 0; i < s.iterations.size() && i < 8; ++i) {
            const IterRecord& r = s.iterations[i];
            double* o = iters + 8 * i;
            o[0] = r.cost; o[1] = r.cost_change; o[2] = r.gradient_max_norm; o[3] = r.step_norm; o[4] = r.relative_decrease; o[5] = r.radius; o[6] = r.valid; o[7] = r.successful;
        }
}

// ---- odometry -------------------------------------------------------------------------------------------
void* orc_odom_create(const orc_config* cfg) { Odometry* o = new Odometry(); o->cfg = to_cfg(cfg); return o; }
void orc_odom_destroy(void* h) { delete (Odometry*)h; }
// pose_out: q_w_curr[4] (xyzw), t_w_curr[3], para_q[4], para_t[3]
void orc_odom_step(void* h, const float* sharp, int ns, const float* less_sharp, int nls, const float* flat, int nf, const float* less_flat,
                   int nlf, double pose_out[14])
{
    Odometry* o = (Odometry*)h;
    o->step(to_cloud(sharp, ns), to_cloud(less_sharp, nls), to_cloud(flat, nf), to_cloud(less_flat, nlf));
    std::memcpy(pose_out, o->q_w_curr, sizeof(double) * 4);
    std::memcpy(pose_out + 4, o->t_w_curr, sizeof(double) * 3);
    std::memcpy(pose_out + 7, o->para_q, sizeof(double) * 4);
    std::memcpy(pose_out + 11, o->para_t, sizeof(double) * 3);
}
void orc_odom_set_warm_start(void* h, const double para_q[4], const double para_t[3])
{
    Odometry* o = (Odometry*)h;
    std::memcpy(o->para_q, para_q, sizeof(double) * 4);
    std::memcpy(o->para_t, para_t, sizeof(double) * 3);
}
// per outer iteration (3): {corner_corr, plane_corr, plane_selected, initial_cost, final_cost, jac_evals, cost_evals, termination}
int orc_odom_stats(void* h, double out[24])
{
    Odometry* o = (Odometry*)h;
    for (size_t i = 0; i < o->last_stats.size() && i < 3; ++i) {
        const OdomIterStats& s = o->last_stats[i];
        double* d = out + 8 * i;
        d[0] = s.corner_corr; d[1] = s.plane_corr; d[2] = s.plane_selected; d[3] = s.solve.initial_cost; d[4] = s.solve.final_cost;
        d[5] = s.solve.num_jacobian_evals; d[6] = s.solve.num_cost_evals; d[7] = s.solve.termination;
    }
    return (int)o->last_stats.size();
}
// association dump of the last outer iteration: corner triples / plane quadruples (caller sizes by the sharp / flat counts)
void orc_odom_assoc(void* h, int* n_corner, int* corner, int* n_plane, int* plane)
{
    Odometry* o = (Odometry*)h;
    *n_corner = (int)o->last_corner_assoc.size();
    *n_plane = (int)o->last_plane_assoc.size();
    if (corner) for (size_t i = 0; i < o->last_corner_assoc.size(); ++i) std::memcpy(corner + 3 * i, o->last_corner_assoc[i].data(), sizeof(int) * 3);
    if (plane) for (size_t i = 0; i < o->last_plane_assoc.size(); ++i) std::memcpy(plane + 4 * i, o->last_plane_assoc[i].data(), sizeof(int) * 4);
}

// ---- mapping --------------------------------------------------------------------------------------------
void* orc_map_create(const orc_config* cfg) { Mapping* m = new Mapping(); m->cfg = to_cfg(cfg); return m; }
void orc_map_destroy(void* h) { delete (Mapping*)h; }
void orc_map_insert(void* h, const float* corner, int nc, const float* surf, int ns) { ((Mapping*)h)->insert_map_points(to_cloud(corner, nc), to_cloud(surf, ns)); }
// pose_out: q_w_curr[4], t_w_curr[3]; info: {rc, map_corner, map_surf, stack_corner, stack_surf, corner_num, surf_num}
int orc_map_step(void* h, const float* corner_last, int nc, const float* surf_last, int ns, const double q_wodom[4], const double t_wodom[3],
                 double pose_out[7], int info[8])
{
    Mapping* m = (Mapping*)h;
    const int rc = m->step(to_cloud(corner_last, nc), to_cloud(surf_last, ns), q_wodom, t_wodom);
    std::memcpy(pose_out, m->parameters, sizeof(double) * 7);
    if (info) { info[0] = rc; info[1] = m->last_map_corner; info[2] = m->last_map_surf; info[3] = m->last_stack_corner; info[4] = m->last_stack_surf; info[5] = m->last_corner_num; info[6] = m->last_surf_num; info[7] = m->last_vote_selected; }
    return rc;
}
long long orc_map_total_points(void* h, int which)
{
    Mapping* m = (Mapping*)h;
    long long n = 0;
    for (const auto& v : (which == 0 ? m->cornerArray : m->surfArray)) n += (long long)v.size();
    return n;
}

// ---- whole pipeline (scanRegistration -> laserOdometry -> laserMapping), timed per stage ---------------
struct Pipeline { Config cfg; Odometry odom; Mapping map; bool with_mapping; };
void* orc_pipeline_create(const orc_config* cfg, int with_mapping)
{
    Pipeline* p = new Pipeline();
    p->cfg = to_cfg(cfg); p->odom.cfg = p->cfg; p->map.cfg = p->cfg; p->with_mapping = with_mapping != 0;
    return p;
}
void orc_pipeline_destroy(void* h) { delete (Pipeline*)h; }
// poses: odom q_w[4], t_w[3], mapped q[4], t[3]; ms: extract, odometry, mapping; counts: n_full, sharp, less_sharp, flat, less_flat
int orc_pipeline_step(void* h, const float* pts, int n, int stride_floats, double poses[14], double ms[3], int counts[5])
{
    Pipeline* p = (Pipeline*)h;
    using clk = std::chrono::steady_clock;
    Features f;
    const auto t0 = clk::now();
    const int rc = extract_features(pts, n, stride_floats, p->cfg, f);
    const auto t1 = clk::now();
    if (rc != 0) return rc;
    p->odom.step(f.sharp, f.less_sharp, f.flat, f.less_flat);
    const auto t2 = clk::now();
    std::memcpy(poses, p->odom.q_w_curr, sizeof(double) * 4);
    std::memcpy(poses + 4, p->odom.t_w_curr, sizeof(double) * 3);
    if (p->with_mapping) {
        // laserOdometry publishes laserCloudCornerLast = this frame's less-sharp cloud, SurfLast = less-flat (LO:882-912)
        // (= f.less_sharp / f.less_flat unless the de-skew's TransformToEnd rewrote them, cfg.distortion == 2)
        p->map.step(p->odom.cornerLast, p->odom.surfLast, p->odom.q_w_curr, p->odom.t_w_curr);
        std::memcpy(poses + 7, p->map.parameters, sizeof(double) * 7);
    } else {
        std::memcpy(poses + 7, poses, sizeof(double) * 7);
    }
    const auto t3 = clk::now();
    if (ms) { ms[0] = std::chrono::duration<double, std::milli>(t1 - t0).count(); ms[1] = std::chrono::duration<double, std::milli>(t2 - t1).count(); ms[2] = std::chrono::duration<double, std::milli>(t3 - t2).count(); }
    if (counts) { counts[0] = (int)f.full.size(); counts[1] = (int)f.sharp.size(); counts[2] = (int)f.less_sharp.size(); counts[3] = (int)f.flat.size(); counts[4] = (int)f.less_flat.size(); }
    return 0;
}

}  // extern "C"
