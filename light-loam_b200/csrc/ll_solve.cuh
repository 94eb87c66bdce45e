// Fused residual + Jacobian + Huber + JtJ / Jtr reduction and the Levenberg-Marquardt controller — the
// replacement for ceres::Solve on this path (LO:475-482, 819-825; LM:1865-1872, 2079-2087) and for the
// autodiff cost functors of lidarFactor.hpp (LidarEdgeFactor LF:9-52, LidarPlaneFactor_modify LF:203-251,
// LidarPlaneNormFactor LF:253-285).
//
// With DISTORTION 0 (LO:23) every live factor is r = f(lp), lp = R(q) cp + t, and composing the ambient
// Jacobian with EigenQuaternionManifold's plus-Jacobian gives d lp / d delta = -2 [R cp]x, d lp / d t = I
// (checked against the oracle's Jet autodiff in tests/).  Column order (dx,dy,dz,tx,ty,tz).
//
// Solver restated: TRUST_REGION + LEVENBERG_MARQUARDT, DENSE_QR replaced by a Cholesky solve of the 6x6
// normal equations (the 28-double reduction 21 + 6 + 1 carries everything the controller needs), Jacobi
// scaling from the first Jacobian, HuberLoss(0.1) with the rho'' <= 0 corrector, radius 1e4, accept if
// rho > 1e-3, max 4 iterations, function / gradient / parameter tolerances 1e-6 / 1e-10 / 1e-8.
//
// One CTA per problem (or several, see LmComm); LM_THREADS threads stride over the residual blocks; the reduction
// order is fixed (lane tree, then warps in order, then CTAs in rank order), so results are run-to-run deterministic.
#pragma once
#include <cooperative_groups.h>
#include <float.h>

#include "ll_ctx.h"
#include "ll_device.cuh"

#define LM_THREADS 512   // most threads a solve CTA may have (a launch may use fewer: the loops follow blockDim.x)
#define LM_NRED 28

#define LM_MAX_GPUS 8
#define LM_MAX_PARTS 16
#define LM_MAX_WORLD (LM_MAX_GPUS * LM_MAX_PARTS)
#define LM_MBOX_DOUBLES 32   // 28 used; one mailbox slot = 32 doubles

// Split solve.  One problem can be evaluated by several CTAs — `nparts` CTAs of this GPU and the CTAs of `gworld`
// GPUs (BASELINE config 5) — each owning a share of the residual blocks.  The 28 doubles are
// all-reduced INSIDE the solve kernel: every CTA stores its partial sums into the mailbox of every GPU (local stores
// or NVLink P2P stores), publishes a sequence flag, spins on the flags of its own GPU's mailbox and sums the
// contributions in rank order — identical bits in every CTA, so all of them take the identical LM step, and neither a
// host round trip nor a separate collective launch sits between two evaluations.
// Mailbox of one context: mbox[lane][parity][LM_MAX_WORLD][32] doubles + flag[lane][LM_MAX_WORLD] (sequence numbers).
struct LmComm {
    double* mbox[LM_MAX_GPUS];                  // mailboxes of all GPUs (own included), device pointers valid on this GPU
    unsigned long long* flag[LM_MAX_GPUS];
    const unsigned long long* seq_in;           // [B] collectives executed so far per problem (same on every rank)
    unsigned long long* seq_out;                // [B] written by part 0 of this launch (the host alternates the two arrays)
    int grank, gworld, nparts;
    unsigned long long timeout_ns;              // a peer that never shows up must not hang the GPU
    int cluster;                                // > 1: the parts of a problem are the CTAs of one thread-block cluster (gworld = 1) and
                                                // exchange their partial sums through distributed shared memory, not the mailbox
};

struct LmCtl {   // trust-region controller state (thread 0 only); kept out of the registers of the evaluation loops
    double H[21], g[6], scale[6], diagonal[6];
    double x_cost, x_norm, radius, decrease_factor, se_cost, model_cost_change, gmax, initial_cost;
    int reuse_diagonal, last_successful, atleast_one, iteration, invalid, term, jac_evals, cost_evals;
};
struct LmShared {
    LmCtl ctl;
    double x[7], cand[7];
    double red[LM_THREADS / 32][LM_NRED];
    double out[LM_NRED];
    double pub[2][LM_NRED];   // cluster mode: this CTA's partial sums as its peers read them (slot = parity of the collective)
    int go, go2;   // decisions of the step phase / of the accept phase (two words: the next step phase may write go while
                   // slower threads still read go2)
    int comm_dead;
};
__device__ __forceinline__ unsigned long long lm_globaltimer()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// One-shot all-reduce (sum) of S.out[first..28) over all CTAs working on problem `b`.  Two mailbox slots (parity of
// the collective counter) suffice: a peer can be at most one collective ahead, because it needs this CTA's
// contribution to complete the next one.
__device__ __forceinline__ void lm_allreduce(LmShared& S, const LmComm& C, int b, int part, unsigned long long seq, int first, LaneState* L)
{
    const int tid = threadIdx.x, W = C.gworld * C.nparts, me = C.grank * C.nparts + part;
    const size_t slot = ((size_t)b * 2 + (seq & 1)) * LM_MAX_WORLD * LM_MBOX_DOUBLES;
    if (tid < LM_NRED && tid >= first) {
        const double v = S.out[tid];
        for (int g = 0; g < C.gworld; ++g) {
            volatile double* dst = C.mbox[g] + slot + (size_t)me * LM_MBOX_DOUBLES;
            dst[tid] = v;
        }
    }
    __syncthreads();
    if (tid < C.gworld) {  // publish: one thread per destination GPU
        __threadfence_system();
        volatile unsigned long long* f = C.flag[tid] + (size_t)b * LM_MAX_WORLD + me;
        *f = seq;
    }
    if (tid < W) {         // wait for contribution `tid` to arrive in THIS GPU's mailbox
        volatile unsigned long long* mine = C.flag[C.grank] + (size_t)b * LM_MAX_WORLD + tid;
        if (!S.comm_dead) {
            const unsigned long long t0 = lm_globaltimer();
            while (*mine < seq) {
                if (lm_globaltimer() - t0 > C.timeout_ns) { S.comm_dead = 1; if (L) L->err = LL_E_NCCL; break; }
                __nanosleep(32);
            }
        }
        __threadfence_system();
    }
    __syncthreads();
    if (tid < LM_NRED && tid >= first) {
        volatile const double* src = C.mbox[C.grank] + slot;
        double v = 0.0;
        for (int r = 0; r < W; ++r) v += src[(size_t)r * LM_MBOX_DOUBLES + tid];  // rank order: same bits in every CTA
        S.out[tid] = v;
    }
    __syncthreads();
}

// The same sum over the CTAs of a thread-block cluster (the single-stream path: few problems, many idle SMs).  Every CTA
// publishes its partial sums in its own shared memory, one cluster barrier later every CTA reads all of them in rank
// order through DSMEM - identical bits everywhere again.  One barrier per collective: a CTA can only overwrite a slot
// two collectives later, and it cannot pass the barrier in between before every peer is done reading.
__device__ __forceinline__ void lm_allreduce_cluster(LmShared& S, int nparts, unsigned long long seq, int first)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cl = cg::this_cluster();
    const int tid = threadIdx.x;
    double* pub = S.pub[seq & 1];
    if (tid < LM_NRED && tid >= first) pub[tid] = S.out[tid];
    cl.sync();
    if (tid < LM_NRED && tid >= first) {
        double t[8], v = 0.0;   // nparts <= 8 (portable cluster size); the remote loads are issued together
#pragma unroll
        for (int r = 0; r < 8; ++r) t[r] = r < nparts ? *cl.map_shared_rank(pub + tid, r) : 0.0;
#pragma unroll
        for (int r = 0; r < 8; ++r) v += t[r];
        S.out[tid] = v;
    }
    __syncthreads();
}

// residual-block record (SoA, stride = cap): [0] type, [1..3] cp, [4..6] p0, [7..9] p1, [10] w
//   type 0 EDGE        p0 = a, p1 = b
//   type 1 PLANE_MODIFY p0 = j, p1 = ljm_norm, w = weight
//   type 2 PLANE_NORM   p0 = unit normal, w = negative_OA_dot_norm, p1.x = 2 when the block counts twice (map vote)
// HuberLoss(0.1): returns rho(s); sq = sqrt(rho'(s)) - the Corrector with rho'' <= 0 scales r and J by it
__device__ __forceinline__ double lm_huber(double s, double& sq)
{
    double rho0 = s, rho1 = 1.0;
    if (s > 0.01) {
        const double rr = sqrt(s);
        rho0 = 2.0 * 0.1 * rr - 0.01;
        rho1 = fmax(DBL_MIN, 0.1 / rr);
    }
    sq = sqrt(rho1);
    return rho0;
}
// one residual row: J = d * [ -2 [R cp]x | I ] scaled by sq; JtJ (upper triangle), Jtr
__device__ __forceinline__ void lm_row(double acc[LM_NRED], double d0, double d1, double d2, double rk, double Rx, double Ry, double Rz, double sq)
{
    double J[6];
    J[0] = (d1 * (-2.0 * Rz) + d2 * (2.0 * Ry)) * sq;
    J[1] = (d0 * (2.0 * Rz) + d2 * (-2.0 * Rx)) * sq;
    J[2] = (d0 * (-2.0 * Ry) + d1 * (2.0 * Rx)) * sq;
    J[3] = d0 * sq; J[4] = d1 * sq; J[5] = d2 * sq;
    rk = rk * sq;
    int q = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
#pragma unroll
        for (int c = a; c < 6; ++c) acc[q++] += J[a] * J[c];
    }
#pragma unroll
    for (int a = 0; a < 6; ++a) acc[21 + a] += J[a] * rk;
}
// ---- DISTORTION 1 (LO:23): the factors interpolate the pose per point, q_s = Identity.slerp(s, q), t_s = s t (LF:23-31,
// LF:227-235), and the closed form above (s = 1) no longer applies.  The slow path evaluates the functors the way
// ceres::AutoDiffCostFunction<.., 4, 3> does: forward-mode duals (ceres/jet.h) through Eigen's slerp and q * v, then the
// ambient quaternion columns times EigenQuaternionManifold's plus-Jacobian.  Only the de-skew mode compiles into the
// kernel that runs it (lm_solve<true>); the reference build's path is untouched.
struct LmJet {
    double a, v[7];
    __device__ LmJet() {}
    __device__ LmJet(double s) : a(s) { for (int i = 0; i < 7; ++i) v[i] = 0.0; }
    __device__ LmJet(double s, int k) : a(s) { for (int i = 0; i < 7; ++i) v[i] = 0.0; v[k] = 1.0; }
};
__device__ __forceinline__ LmJet operator+(const LmJet& f, const LmJet& g) { LmJet h; h.a = f.a + g.a; for (int i = 0; i < 7; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
__device__ __forceinline__ LmJet operator-(const LmJet& f, const LmJet& g) { LmJet h; h.a = f.a - g.a; for (int i = 0; i < 7; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
__device__ __forceinline__ LmJet operator-(const LmJet& f) { LmJet h; h.a = -f.a; for (int i = 0; i < 7; ++i) h.v[i] = -f.v[i]; return h; }
__device__ __forceinline__ LmJet operator*(const LmJet& f, const LmJet& g) { LmJet h; h.a = f.a * g.a; for (int i = 0; i < 7; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
__device__ __forceinline__ LmJet operator/(const LmJet& f, const LmJet& g)
{
    LmJet h;
    const double gi = 1.0 / g.a, fg = f.a * gi;
    h.a = fg;
    for (int i = 0; i < 7; ++i) h.v[i] = (f.v[i] - fg * g.v[i]) * gi;
    return h;
}
__device__ __forceinline__ bool operator<(const LmJet& f, const LmJet& g) { return f.a < g.a; }
__device__ __forceinline__ bool operator>=(const LmJet& f, const LmJet& g) { return f.a >= g.a; }
__device__ __forceinline__ LmJet lm_sqrt(const LmJet& f) { LmJet h; const double t = sqrt(f.a); h.a = t; const double k = 1.0 / (2.0 * t); for (int i = 0; i < 7; ++i) h.v[i] = f.v[i] * k; return h; }
__device__ __forceinline__ LmJet lm_sin(const LmJet& f) { LmJet h; double sn, cs; sincos(f.a, &sn, &cs); h.a = sn; for (int i = 0; i < 7; ++i) h.v[i] = cs * f.v[i]; return h; }
__device__ __forceinline__ LmJet lm_acos(const LmJet& f) { LmJet h; h.a = acos(f.a); const double t = -1.0 / sqrt(1.0 - f.a * f.a); for (int i = 0; i < 7; ++i) h.v[i] = t * f.v[i]; return h; }
__device__ __forceinline__ LmJet lm_abs(const LmJet& f) { return f.a < 0.0 ? -f : f; }
__device__ __forceinline__ double lm_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ double lm_sin(double x) { return sin(x); }
__device__ __forceinline__ double lm_acos(double x) { return acos(x); }
__device__ __forceinline__ double lm_abs(double x) { return fabs(x); }
// Eigen QuaternionBase::slerp(t, other) with *this = Identity (LF:25-26, LO:86): d = this.dot(other)
template <typename T>
__device__ __forceinline__ void lm_identity_slerp(const T& t, const T q[4], T out[4])
{
    const T one = T(1.0 - 2.220446049250313e-16);
    const T d = T(0.0) * q[0] + T(0.0) * q[1] + T(0.0) * q[2] + T(1.0) * q[3];
    const T absD = lm_abs(d);
    T scale0, scale1;
    if (absD >= one) {
        scale0 = T(1.0) - t;
        scale1 = t;
    } else {
        const T theta = lm_acos(absD);
        const T sinTheta = lm_sin(theta);
        scale0 = lm_sin((T(1.0) - t) * theta) / sinTheta;
        scale1 = lm_sin(t * theta) / sinTheta;
    }
    if (d < T(0.0)) scale1 = -scale1;
    out[0] = scale0 * T(0.0) + scale1 * q[0];
    out[1] = scale0 * T(0.0) + scale1 * q[1];
    out[2] = scale0 * T(0.0) + scale1 * q[2];
    out[3] = scale0 * T(1.0) + scale1 * q[3];
}
// LidarEdgeFactor (LF:23-38) / LidarPlaneFactor_modify (LF:227-237) for T = double or LmJet; returns the residual count
template <typename T>
__device__ __forceinline__ int lm_functor_s(int type, const double* rec /* cp, p0, p1, w */, double s, const T q[4], const T t[3], T res[3])
{
    T qs[4];
    lm_identity_slerp(T(s), q, qs);
    const T ts[3] = {T(s) * t[0], T(s) * t[1], T(s) * t[2]};
    const T cp[3] = {T(rec[0]), T(rec[1]), T(rec[2])};
    // Eigen _transformVector: uv = 2 (u x v); v + w uv + u x uv
    T uv[3] = {qs[1] * cp[2] - qs[2] * cp[1], qs[2] * cp[0] - qs[0] * cp[2], qs[0] * cp[1] - qs[1] * cp[0]};
    uv[0] = uv[0] + uv[0]; uv[1] = uv[1] + uv[1]; uv[2] = uv[2] + uv[2];
    const T lp[3] = {(cp[0] + qs[3] * uv[0]) + (qs[1] * uv[2] - qs[2] * uv[1]) + ts[0], (cp[1] + qs[3] * uv[1]) + (qs[2] * uv[0] - qs[0] * uv[2]) + ts[1],
                     (cp[2] + qs[3] * uv[2]) + (qs[0] * uv[1] - qs[1] * uv[0]) + ts[2]};
    if (type == 0) {
        const T a[3] = {T(rec[3]), T(rec[4]), T(rec[5])}, b[3] = {T(rec[6]), T(rec[7]), T(rec[8])};
        const T u[3] = {lp[0] - a[0], lp[1] - a[1], lp[2] - a[2]}, v[3] = {lp[0] - b[0], lp[1] - b[1], lp[2] - b[2]};
        const T de[3] = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
        const T dn = lm_sqrt(de[0] * de[0] + de[1] * de[1] + de[2] * de[2]);
        res[0] = (u[1] * v[2] - u[2] * v[1]) / dn;
        res[1] = (u[2] * v[0] - u[0] * v[2]) / dn;
        res[2] = (u[0] * v[1] - u[1] * v[0]) / dn;
        return 3;
    }
    const T j[3] = {T(rec[3]), T(rec[4]), T(rec[5])}, n[3] = {T(rec[6]), T(rec[7]), T(rec[8])};
    res[0] = ((lp[0] - j[0]) * n[0] + (lp[1] - j[1]) * n[1] + (lp[2] - j[2]) * n[2]) * T(rec[9]);
    return 1;
}
// one block of the de-skew mode: cost, JtJ and Jtr into acc
static __device__ __noinline__ void lm_block_distort(int type, const double* rec, double s, const double* x, double acc[LM_NRED])
{
    LmJet q[4], t[3], r[3];
    for (int k = 0; k < 4; ++k) q[k] = LmJet(x[k], k);
    for (int k = 0; k < 3; ++k) t[k] = LmJet(x[4 + k], 4 + k);
    const int nr = lm_functor_s<LmJet>(type, rec, s, q, t, r);
    double ss = 0.0;
    for (int k = 0; k < nr; ++k) ss += r[k].a * r[k].a;
    double sq;
    acc[27] += 0.5 * lm_huber(ss, sq);
    // EigenQuaternionManifold::PlusJacobian, rows x, y, z, w
    const double PJ[4][3] = {{x[3], x[2], -x[1]}, {-x[2], x[3], x[0]}, {x[1], -x[0], x[3]}, {-x[0], -x[1], -x[2]}};
    for (int k = 0; k < nr; ++k) {
        double J[6];
        for (int c = 0; c < 3; ++c) J[c] = (r[k].v[0] * PJ[0][c] + r[k].v[1] * PJ[1][c] + r[k].v[2] * PJ[2][c] + r[k].v[3] * PJ[3][c]) * sq;
        for (int c = 0; c < 3; ++c) J[3 + c] = r[k].v[4 + c] * sq;
        const double rk = r[k].a * sq;
        int qi = 0;
        for (int a = 0; a < 6; ++a)
            for (int c = a; c < 6; ++c) acc[qi++] += J[a] * J[c];
        for (int a = 0; a < 6; ++a) acc[21 + a] += J[a] * rk;
    }
}

struct LmRecord { double v[12]; };
__device__ __forceinline__ void lm_load_record(LmRecord& r, const double* __restrict__ blk, int cap, int i)
{
#pragma unroll
    for (int k = 0; k < 11; ++k) r.v[k] = __ldg(blk + (size_t)k * cap + i);
}
template <bool DIST = false>
__device__ __forceinline__ void lm_accumulate(const double* __restrict__ blk, int cap, int nb, const double* x, double acc[LM_NRED], int part = 0, int nparts = 1)
{
#pragma unroll
    for (int k = 0; k < LM_NRED; ++k) acc[k] = 0.0;
    // the record of the thread's next block is in flight while the current one is evaluated (all eleven fields are
    // loaded before the type is looked at: a slot without a correspondence still holds readable numbers)
    const int stride = (int)blockDim.x * nparts;   // blockDim.x <= LM_THREADS
    int i = part * (int)blockDim.x + threadIdx.x;
    LmRecord cur, nxt;
    if (i < nb) lm_load_record(cur, blk, cap, i);
    for (; i < nb; i += stride) {
        if (i + stride < nb) lm_load_record(nxt, blk, cap, i + stride);
        const int type = (int)cur.v[0];
        if (DIST && (type == 0 || type == 1)) {   // de-skew mode: per-point interpolation ratio s in record slot 11
            lm_block_distort(type, &cur.v[1], __ldg(blk + (size_t)11 * cap + i), x, acc);
        } else
        if (type >= 0) {  // dense mapping records: type -1 = slot without a correspondence
        const double cpx = cur.v[1], cpy = cur.v[2], cpz = cur.v[3];
        const double ax = cur.v[4], ay = cur.v[5], az = cur.v[6];
        const double bx = cur.v[7], by = cur.v[8], bz = cur.v[9];
        const double w = cur.v[10];
        double Rx, Ry, Rz;
        quat_rotate(x, cpx, cpy, cpz, Rx, Ry, Rz);
        const double lx = Rx + x[4], ly = Ry + x[5], lz = Rz + x[6];
        if (type == 0) {
            // LidarEdgeFactor: r = ((lp - a) x (lp - b)) / |a - b| ; d r / d lp = [b - a]x / |a - b|
            const double ux = lx - ax, uy = ly - ay, uz = lz - az, vx = lx - bx, vy = ly - by, vz = lz - bz;
            const double dex = ax - bx, dey = ay - by, dez = az - bz;
            const double dn = sqrt(dex * dex + dey * dey + dez * dez);
            const double r0 = (uy * vz - uz * vy) / dn, r1 = (uz * vx - ux * vz) / dn, r2 = (ux * vy - uy * vx) / dn;
            const double ex = -dex / dn, ey = -dey / dn, ez = -dez / dn;
            const double s = (r0 * r0 + r1 * r1) + r2 * r2;
            double sq;
            acc[27] += 0.5 * lm_huber(s, sq);
            lm_row(acc, 0.0, -ez, ey, r0, Rx, Ry, Rz, sq);
            lm_row(acc, ez, 0.0, -ex, r1, Rx, Ry, Rz, sq);
            lm_row(acc, -ey, ex, 0.0, r2, Rx, Ry, Rz, sq);
        } else {
            // type 1 LidarPlaneFactor_modify: r = w (lp - j) . n ; type 2 LidarPlaneNormFactor: r = n . lp + d
            const double r0 = type == 1 ? ((lx - ax) * bx + (ly - ay) * by + (lz - az) * bz) * w : (ax * lx + ay * ly + az * lz) + w;
            const double d0 = type == 1 ? w * bx : ax, d1 = type == 1 ? w * by : ay, d2 = type == 1 ? w * bz : az;
            // a PLANE_NORM record the scan-to-map vote selected stands for two identical blocks (LM:2064-2067): p1.x = 2
            const bool twice = type == 2 && bx == 2.0;
            double sq;
            const double rho = lm_huber(0.0 + r0 * r0, sq);
            acc[27] += twice ? rho : 0.5 * rho;
            lm_row(acc, d0, d1, d2, r0, Rx, Ry, Rz, twice ? sq * 1.4142135623730951 : sq);
        }
        }
        cur = nxt;
    }
}

// fixed-order block reduction of acc[first..LM_NRED) into S.out.
// Warp level: a transposed reduction - every round the lanes swap half of the values they still hold with the lane
// `m` away and add, so after rounds m = 16, 8, 4, 2, 1 each lane owns the warp total of ONE value (31 exchanges of
// a double instead of 5 per value).  The tree is fixed, so the sums are run-to-run deterministic.
template <int CNT, int M>
__device__ __forceinline__ void lm_fold(double (&v)[32], int lane)
{
    const bool upper = (lane & M) != 0;
#pragma unroll
    for (int k = 0; k < CNT / 2; ++k) {
        const double send = upper ? v[k] : v[k + CNT / 2];
        const double keep = upper ? v[k + CNT / 2] : v[k];
        v[k] = keep + shfl_xor_f64(send, M);
    }
}
__device__ __forceinline__ void lm_reduce(LmShared& S, double acc[LM_NRED])
{
    const int lane = lane_id(), w = warp_id();
    double v[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) v[k] = k < LM_NRED ? acc[k] : 0.0;
    lm_fold<32, 16>(v, lane);
    lm_fold<16, 8>(v, lane);
    lm_fold<8, 4>(v, lane);
    lm_fold<4, 2>(v, lane);
    lm_fold<2, 1>(v, lane);
    // the lane's value index: bit 4 of the lane picked the upper half of 32, bit 3 of 16, ... bit 0 of 2
    const int idx = ((lane >> 4) & 1) * 16 + ((lane >> 3) & 1) * 8 + ((lane >> 2) & 1) * 4 + ((lane >> 1) & 1) * 2 + (lane & 1);
    if (idx < LM_NRED) S.red[w][idx] = v[0];
    __syncthreads();
    if (threadIdx.x < LM_NRED) {
        double t = 0.0;
        const int nw = (int)blockDim.x >> 5;
#pragma unroll 4
        for (int ww = 0; ww < nw; ++ww) t += S.red[ww][threadIdx.x];
        S.out[threadIdx.x] = t;
    }
    __syncthreads();
}

// EigenQuaternionManifold::Plus on q, Euclidean on t
__device__ __forceinline__ void lm_plus(const double* x, const double* delta, double* out)
{
    const double nd = sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
    if (nd == 0.0) {
        for (int i = 0; i < 4; ++i) out[i] = x[i];
    } else {
        double sn, cs;
        sincos(nd, &sn, &cs);
        const double sbd = sn / nd;
        const double dq[4] = {sbd * delta[0], sbd * delta[1], sbd * delta[2], cs};
        quat_mul(dq, x, out);
    }
    for (int i = 0; i < 3; ++i) out[4 + i] = x[4 + i] + delta[3 + i];
}

// 6x6 SPD solve A y = b by Cholesky (A given as packed upper triangle row-major, 21 entries). false if not SPD.
// One thread runs it between two evaluations while the CTA waits, so the dependent chain is what counts: the
// pivots are kept as reciprocal square roots (one rsqrt per column, multiplications everywhere else).
__device__ __forceinline__ bool chol6_solve(const double* Ap, const double* b, double* y)
{
    double Lm[6][6], inv[6];
    int q = 0;
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int c = a; c < 6; ++c) { Lm[a][c] = Ap[q]; Lm[c][a] = Ap[q]; ++q; }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double d = Lm[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) d -= Lm[j][k] * Lm[j][k];
        if (!(d > 0.0)) return false;
        inv[j] = rsqrt(d);
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            double s = Lm[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) s -= Lm[i][k] * Lm[j][k];
            Lm[i][j] = s * inv[j];
        }
    }
    double z[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double s = b[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s -= Lm[i][k] * z[k];
        z[i] = s * inv[i];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double s = z[i];
#pragma unroll
        for (int k = i + 1; k < 6; ++k) s -= Lm[k][i] * y[k];
        y[i] = s * inv[i];
        if (!isfinite(y[i])) return false;
    }
    return true;
}

// The whole Solve. q_io / t_io point at the parameter blocks (global memory); all threads of the CTA call it.
template <bool DIST = false>
static __device__ __noinline__ void lm_solve(const double* blk, int cap, int nb, double* q_io, double* t_io, LaneState* L, int slot,
                                             const LmComm* comm = nullptr, int comm_b = 0, int part = 0)
{
    __shared__ LmShared S;
    const int tid = threadIdx.x;
    const bool clus = comm != nullptr && comm->cluster > 1;   // parts = CTAs of a cluster: same lane, same nb, DSMEM exchange
    const bool dist = comm != nullptr && !clus && comm->gworld * comm->nparts > 1;
    const int nparts = clus ? comm->cluster : (dist ? comm->nparts : 1);
    unsigned long long seq = dist ? comm->seq_in[comm_b] : 0ull;   // collectives so far; every thread keeps its own copy
    if (tid == 0) S.comm_dead = 0;
    if (tid < 4) S.x[tid] = q_io[tid];
    if (tid >= 4 && tid < 7) S.x[tid] = t_io[tid - 4];
    __syncthreads();
    if (dist) {  // the ranks agree on the problem size first: "nothing to minimise" must be a collective decision
        if (tid < LM_NRED) S.out[tid] = tid == 27 ? (double)(nb > 0 && part == 0 ? nb : 0) : 0.0;
        __syncthreads();
        lm_allreduce(S, *comm, comm_b, part, ++seq, 27, L);
        const int nb_all = (int)S.out[27];
        __syncthreads();
        if (nb_all <= 0) nb = 0; else if (nb <= 0) nb = -1;  // -1: this GPU owns no block but takes part in every reduction
    }
    if (nb == 0 || (!dist && nb < 0)) {  // nothing to minimise: parameters untouched
        if (tid == 0 && L) { L->initial_cost[slot] = 0; L->final_cost[slot] = 0; L->jac_evals[slot] = 0; L->cost_evals[slot] = 0; L->termination[slot] = -1; }
        if (tid == 0 && dist && part == 0) comm->seq_out[comm_b] = seq;
        return;
    }
    double acc[LM_NRED];
    LmCtl& C = S.ctl;
#ifdef LL_LM_TIMING   // development build: cycles per phase of the solve, summed into LaneState::dbg (LL_DEBUG_LM prints them)
    long long tmark = clock64();
#define LMT(slot) do { if (tid == 0 && part == 0 && L) { const long long now_ = clock64(); atomicAdd(&L->dbg[slot], (int)(now_ - tmark)); tmark = now_; } } while (0)
#else
#define LMT(slot) ((void)0)
#endif
    if (tid == 0) {
        C.x_cost = 0; C.x_norm = 0; C.radius = 1e4; C.decrease_factor = 2.0; C.se_cost = 0; C.model_cost_change = 0; C.gmax = 0; C.initial_cost = 0;
        C.reuse_diagonal = 0; C.last_successful = 1; C.atleast_one = 0; C.iteration = 0; C.invalid = 0; C.term = 0; C.jac_evals = 0; C.cost_evals = 0;
    }
    auto norm7 = [](const double* v) { double s = 0; for (int i = 0; i < 7; ++i) s += v[i] * v[i]; return sqrt(s); };

    // IterationZero: cost, gradient, Jacobian (as JtJ) at x
    lm_accumulate<DIST>(blk, cap, nb, S.x, acc, part, nparts);
    LMT(0);
    lm_reduce(S, acc);
    if (dist) lm_allreduce(S, *comm, comm_b, part, ++seq, 0, L);
    if (clus) lm_allreduce_cluster(S, nparts, ++seq, 0);
    LMT(1);
    if (tid == 0) {
        for (int k = 0; k < 21; ++k) C.H[k] = S.out[k];
        for (int k = 0; k < 6; ++k) C.g[k] = S.out[21 + k];
        C.x_cost = S.out[27];
        C.initial_cost = C.x_cost;
        C.se_cost = C.x_cost;
        C.jac_evals = 1;
        C.x_norm = norm7(S.x);
        const int dq[6] = {0, 6, 11, 15, 18, 20};  // packed index of the C.diagonal
        for (int c = 0; c < 6; ++c) C.scale[c] = 1.0 / (1.0 + sqrt(C.H[dq[c]]));  // jacobi_scaling, first Jacobian only
        double ng[6], xp[7];
        for (int c = 0; c < 6; ++c) ng[c] = -C.g[c];
        lm_plus(S.x, ng, xp);
        C.gmax = 0;
        for (int i = 0; i < 7; ++i) C.gmax = fmax(C.gmax, fabs(S.x[i] - xp[i]));
    }
    for (;;) {
        // ---- thread 0: finalize previous C.iteration, compute the trust-region step ----------------------
        if (tid == 0) {
            int go = 1;  // 1: evaluate candidate cost, 0: stop
            for (;;) {
                if (C.iteration >= 4) { C.term = 0; go = 0; break; }                         // max_num_iterations (LO:822)
                if (C.last_successful && C.gmax <= 1e-10) { C.term = 1; go = 0; break; }       // gradient_tolerance
                if (C.radius <= 1e-32) { C.term = 4; go = 0; break; }                        // min_trust_region_radius
                ++C.iteration;
                const int dq[6] = {0, 6, 11, 15, 18, 20};
                if (!C.reuse_diagonal)
                    for (int c = 0; c < 6; ++c) C.diagonal[c] = fmin(fmax(C.H[dq[c]] * C.scale[c] * C.scale[c], 1e-6), 1e32);
                // (S C.H S + D^2) y = S C.g ; step = -y
                double A[21], rhs[6], y[6], step[6];
                int q = 0;
                for (int a = 0; a < 6; ++a)
                    for (int c = a; c < 6; ++c) { A[q] = C.H[q] * C.scale[a] * C.scale[c]; ++q; }
                for (int c = 0; c < 6; ++c) { A[dq[c]] += C.diagonal[c] / C.radius; rhs[c] = C.g[c] * C.scale[c]; }
                const bool ok = chol6_solve(A, rhs, y);
                C.reuse_diagonal = true;
                bool valid = false;
                if (ok) {
                    for (int c = 0; c < 6; ++c) step[c] = -y[c];
                    // C.model_cost_change = -(J step).(r + J step / 2) = -step.(S C.g) - step^T (S C.H S) step / 2
                    double lin = 0, quad = 0;
                    for (int c = 0; c < 6; ++c) lin += step[c] * rhs[c];
                    q = 0;
                    for (int a = 0; a < 6; ++a)
                        for (int c = a; c < 6; ++c) {
                            const double hv = C.H[q] * C.scale[a] * C.scale[c];
                            quad += (a == c ? 1.0 : 2.0) * hv * step[a] * step[c];
                            ++q;
                        }
                    C.model_cost_change = -lin - 0.5 * quad;
                    valid = C.model_cost_change > 0.0;
                }
                if (!valid) {  // HandleInvalidStep
                    C.last_successful = false;
                    if (++C.invalid >= 5) { C.term = 5; go = 0; break; }
                    C.radius *= 0.5;
                    continue;
                }
                C.invalid = 0;
                double delta[6];
                for (int c = 0; c < 6; ++c) delta[c] = step[c] * C.scale[c];
                lm_plus(S.x, delta, S.cand);
                break;
            }
            S.go = go;
        }
        __syncthreads();
        LMT(2);
        if (!S.go) break;
        // ---- all threads: cost at the candidate - and its linearisation in the same pass.  Ceres evaluates the cost
        // first and the Jacobian only once the step is accepted; accepted steps are the rule (and a cost-only pass costs
        // nearly as much as a full one: the records dominate), so the Jacobian sums are taken speculatively and simply
        // dropped when the step is rejected.  Same numbers, one pass and one reduction per iteration instead of two.
        lm_accumulate<DIST>(blk, cap, nb, S.cand, acc, part, nparts);
        LMT(3);
        lm_reduce(S, acc);
        if (dist) lm_allreduce(S, *comm, comm_b, part, ++seq, 0, L);
        if (clus) lm_allreduce_cluster(S, nparts, ++seq, 0);
        LMT(4);
        if (tid == 0) {
            const double cand_cost = S.out[27];
            ++C.cost_evals;
            int go = 2;  // 2: accepted, 1: rejected -> next step, 0: stop
            double sn = 0;
            for (int i = 0; i < 7; ++i) sn += (S.x[i] - S.cand[i]) * (S.x[i] - S.cand[i]);
            if (C.atleast_one && sqrt(sn) <= 1e-8 * (C.x_norm + 1e-8)) { C.term = 2; go = 0; }           // parameter_tolerance
            else if (C.atleast_one && fabs(C.x_cost - cand_cost) <= 1e-6 * C.x_cost) { C.term = 3; go = 0; }  // function_tolerance
            else {
                const double rel = (C.se_cost - cand_cost) / C.model_cost_change;
                if (rel > 1e-3) {  // min_relative_decrease: HandleSuccessfulStep
                    for (int i = 0; i < 7; ++i) S.x[i] = S.cand[i];
                    C.x_norm = norm7(S.x);
                    { const double u = 2.0 * rel - 1.0; C.radius = C.radius / fmax(1.0 / 3.0, 1.0 - u * u * u); }
                    C.radius = fmin(1e16, C.radius);
                    C.decrease_factor = 2.0;
                    C.reuse_diagonal = false;
                    C.se_cost = cand_cost;
                    C.atleast_one = true;
                    C.last_successful = true;
                    go = 2;
                    // EvaluateGradientAndJacobian at the accepted point: the sums are already here
                    for (int k = 0; k < 21; ++k) C.H[k] = S.out[k];
                    for (int k = 0; k < 6; ++k) C.g[k] = S.out[21 + k];
                    C.x_cost = cand_cost;
                    ++C.jac_evals;
                    double ng[6], xp[7];
                    for (int c = 0; c < 6; ++c) ng[c] = -C.g[c];
                    lm_plus(S.x, ng, xp);
                    C.gmax = 0;
                    for (int i = 0; i < 7; ++i) C.gmax = fmax(C.gmax, fabs(S.x[i] - xp[i]));
                } else {  // StepRejected
                    C.radius = C.radius / C.decrease_factor;
                    C.decrease_factor *= 2.0;
                    C.reuse_diagonal = true;
                    C.last_successful = false;
                    go = 1;
                }
            }
            S.go2 = go;
        }
        __syncthreads();
        LMT(5);
        if (S.go2 == 0) break;
    }
    if (clus) cooperative_groups::this_cluster().sync();   // no CTA leaves while a peer may still be reading its shared memory
    if (tid == 0 && part == 0) {  // every part holds the same result; one writes it
        if (!S.comm_dead) {       // a collective that timed out leaves partial sums behind: the parameters stay untouched (L->err = LL_E_NCCL)
            for (int i = 0; i < 4; ++i) q_io[i] = S.x[i];
            for (int i = 0; i < 3; ++i) t_io[i] = S.x[4 + i];
        }
        if (dist) comm->seq_out[comm_b] = seq;
        if (L) {
            L->initial_cost[slot] = C.initial_cost;
            L->final_cost[slot] = C.x_cost;
            L->jac_evals[slot] = C.jac_evals;
            L->cost_evals[slot] = C.cost_evals;
            L->termination[slot] = C.term;
        }
    }
}

// host side: launch `kern` on grid (lanes, parts) with the parts of a lane forming one thread-block cluster
template <typename... KArgs, typename... Args>
static inline cudaError_t lm_launch_cluster(void (*kern)(KArgs...), int n_lanes, int parts, int threads, cudaStream_t s, Args... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_lanes, parts); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = 0; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = (unsigned)parts; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, args...);
}
// Can the device schedule a cluster of `parts` CTAs of `kern` at all (enough SMs per GPC, MIG slices, ...)?  Asked once per
// kernel and context; a refusal makes the caller fall back to one CTA per problem instead of failing the launch.
template <typename... KArgs>
static inline bool lm_cluster_fits(void (*kern)(KArgs...), int parts, int threads)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(1, parts); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = 0;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = (unsigned)parts; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) { cudaGetLastError(); return false; }
    return n >= 1;
}
// CTAs per cluster for a solve over n_lanes problems
// (LL_LM_CLUSTER overrides; 1 = off)
static inline int lm_cluster_size(int n_lanes, int n_sm)
{
    // measured on B200 (256-thread to 512-thread CTAs, ~2000 blocks per problem): 8 or 4 CTAs pay while half the SMs stay
    // free for the other problems' CTAs, 2 CTAs as long as every CTA has an SM
    int p = 1;
    if (n_lanes * 8 <= n_sm / 2) p = 8;
    else if (n_lanes * 4 <= n_sm / 2) p = 4;
    else if (n_lanes * 2 <= n_sm) p = 2;
    if (const char* e = getenv("LL_LM_CLUSTER")) { const int v = atoi(e); if ((v == 1 || v == 2 || v == 4 || v == 8) && v * n_lanes <= n_sm) p = v; }
    return p;
}

