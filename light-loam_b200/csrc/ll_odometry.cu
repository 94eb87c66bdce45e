// Scan-to-scan odometry on the device — replaces laserOdometry.cpp:425-896.
//
//   k_grid_count / k_grid_scan / k_grid_scatter   kdtree*Last->setInputCloud (LO:895-896) -> hashed grid
//   k_odom_assoc     LO:491-556 + LO:653-723: TransformToStart, 1-NN (d2 < 25), ring-window 2nd / 3rd point;
//                    one warp per feature point, literal scan-loop semantics evaluated 32 candidates at a time
//   k_odom_prep      order-preserving compaction of the matches, graph_based_correspondence_vote_simple
//                    (LO:165-342) for planes when now_frame > 5, residual-block records (LF ctor maths)
//   k_lm_solve       ceres::Solve as configured at LO:819-825 / LM:2079-2087: Levenberg-Marquardt on the
//                    6-dim tangent space with Huber(0.1); residual + analytic Jacobian + JtJ / Jtr / cost in one
//                    pass (fixed-order warp-shuffle + smem reduction of 28 doubles), LM controller on-device
//   k_odom_finalize  pose accumulation LO:830-831, cloud swap LO:882-891, frame counters
//
// All decisions (accept / reject, radius, termination) run on the device: one launch per Solve, no host sync.
#include <limits.h>
#include <math.h>
#include <stdlib.h>

#include "ll_ctx.h"
#include "ll_device.cuh"
#include "ll_knn.cuh"
#include "ll_solve.cuh"

namespace {

// ------------------------------------------------------------------------------------------------------
// grid build
// ------------------------------------------------------------------------------------------------------
struct GridSrc {           // where a lane's points and count come from
    const float4* pts[2];  // ping-pong buffers (or both the same)
    size_t lane_stride;    // points per lane
    int which;             // 0: (n_less_sharp, slot cur) 1: (n_less_flat, slot cur) 2: (n_map_corner, slot 0) 3: (n_map_surf, slot 0)
    int az_bins;           // 0: spatial hash grid; > 0: ring x azimuth-bin index with this many bins per ring
    int rings;
};
__device__ __forceinline__ int src_bucket(const GridSrc& S, const float4 p, int T, float inv_h)
{
    if (S.az_bins == 0) return cell_bucket((int)floorf(p.x * inv_h), (int)floorf(p.y * inv_h), (int)floorf(p.z * inv_h), T - 1);
    int r = (int)p.w;
    r = r < 0 ? 0 : (r >= S.rings ? S.rings - 1 : r);
    return r * S.az_bins + azimuth_bin(p.x, p.y, S.az_bins);
}
__device__ __forceinline__ int grid_src_count(const GridSrc& S, const LaneState& L)
{
    switch (S.which) {
        case 0: return L.n_less_sharp;
        case 1: return L.n_less_flat;
        case 2: return L.n_map_corner;
        default: return L.n_map_surf;
    }
}
__device__ __forceinline__ const float4* grid_src_pts(const GridSrc& S, const LaneState& L, int b)
{
    const int slot = S.which < 2 ? L.cur : 0;
    return S.pts[slot] + (size_t)b * S.lane_stride;
}

__global__ void k_grid_count(GridSrc S, LaneState* lane, int* cursor, int T, float inv_h)
{
    const int b = blockIdx.y;
    LaneState& L = lane[b];
    const int n = grid_src_count(S, L);
    const float4* pts = grid_src_pts(S, L, b);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = pts[i];
        const int bk = src_bucket(S, p, T, inv_h);
        atomicAdd(&cursor[(size_t)b * T + bk], 1);
        if (S.which < 2 && S.az_bins == 0) {  // is the cloud ring-monotone? (LO:504-553 assumes it; the fast association path needs it)
            const int r0 = (int)p.w, r1 = i + 1 < n ? (int)pts[i + 1].w : r0;
            if (r0 < 0 || r0 >= S.rings || r1 < r0) { if (S.which == 0) L.mono_corner = 0; else L.mono_surf = 0; }
        }
    }
}
// two-level exclusive scan of the bucket counts: chunk sums, then per-chunk scan with its base
#define GRID_CHUNK 2048
__global__ void __launch_bounds__(256) k_grid_partial(const int* cursor, int* partial, int T)
{
    __shared__ int ws[40];
    const int b = blockIdx.y, nchunk = T / GRID_CHUNK;
    const int* cur = cursor + (size_t)b * T + (size_t)blockIdx.x * GRID_CHUNK;
    int s = 0;
#pragma unroll
    for (int k = 0; k < GRID_CHUNK / 256; ++k) s += cur[threadIdx.x * (GRID_CHUNK / 256) + k];
    int tot = 0;
    block_exclusive_scan(s, ws, &tot);
    if (threadIdx.x == 0) partial[(size_t)b * (nchunk + 1) + blockIdx.x] = tot;
}
__global__ void __launch_bounds__(256) k_grid_scan(int* cursor, int* start, const int* partial, int T)
{
    __shared__ int ws[40];
    __shared__ int base_s;
    const int b = blockIdx.y, nchunk = T / GRID_CHUNK, chunk = blockIdx.x;
    if (threadIdx.x < 32) {
        int v = 0;
        for (int q = threadIdx.x; q < chunk; q += 32) v += partial[(size_t)b * (nchunk + 1) + q];
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(LL_FULL_MASK, v, d);
        if (threadIdx.x == 0) base_s = v;
    }
    int* cur = cursor + (size_t)b * T + (size_t)chunk * GRID_CHUNK;
    int* st = start + (size_t)b * (T + 1) + (size_t)chunk * GRID_CHUNK;
    const int per = GRID_CHUNK / 256, i0 = threadIdx.x * per;
    int s = 0;
#pragma unroll
    for (int k = 0; k < per; ++k) s += cur[i0 + k];
    int tot = 0;
    int run = block_exclusive_scan(s, ws, &tot) + base_s;  // the scan's barriers also publish base_s
#pragma unroll
    for (int k = 0; k < per; ++k) { const int c = cur[i0 + k]; st[i0 + k] = run; cur[i0 + k] = run; run += c; }
    if (chunk == nchunk - 1 && threadIdx.x == 255) start[(size_t)b * (T + 1) + T] = run;
}
__global__ void k_grid_scatter(GridSrc S, const LaneState* lane, int* cursor, float4* sorted, int T, int cap, float inv_h)
{
    const int b = blockIdx.y;
    const LaneState& L = lane[b];
    const int n = grid_src_count(S, L);
    const float4* pts = grid_src_pts(S, L, b);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = pts[i];
        const int bk = src_bucket(S, p, T, inv_h);
        const int pos = atomicAdd(&cursor[(size_t)b * T + bk], 1);
        int ring = S.which < 2 ? (int)p.w : 0;
        ring = ring < 0 ? 0 : (ring > 255 ? 255 : ring);
        sorted[(size_t)b * cap + pos] = make_float4(p.x, p.y, p.z, __int_as_float((i & 0xFFFFFF) | (ring << 24)));
    }
}

// ---- fused build of the four odometry indexes (corner / surf x spatial hash / ring-azimuth) -------------------
struct IndexSet {
    const float4* pts[2][2];   // [cloud][ping-pong slot]
    size_t lane_stride[2];
    int* cursor[4];            // table = cloud + 2 * kind   (kind 0 spatial, 1 ring-azimuth)
    int* start[4];
    int* partial[4];
    float4* sorted[4];
    int T[4], cap[4];
    int chunk_begin[5];        // prefix of T / GRID_CHUNK over the tables
    float inv_h;
    int az_bins[2];
    int rings;
};
__device__ __forceinline__ int index_bucket(const IndexSet& S, int cloud, int kind, const float4 p)
{
    if (kind == 0) return cell_bucket((int)floorf(p.x * S.inv_h), (int)floorf(p.y * S.inv_h), (int)floorf(p.z * S.inv_h), S.T[cloud] - 1);
    int r = (int)p.w;
    r = r < 0 ? 0 : (r >= S.rings ? S.rings - 1 : r);
    return r * S.az_bins[cloud] + azimuth_bin(p.x, p.y, S.az_bins[cloud]);
}
__global__ void k_index_count(IndexSet S, LaneState* lane)
{
    const int b = blockIdx.y, cloud = blockIdx.z;
    LaneState& L = lane[b];
    const int n = cloud == 0 ? L.n_less_sharp : L.n_less_flat;
    const float4* pts = S.pts[cloud][L.cur] + (size_t)b * S.lane_stride[cloud];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = pts[i];
        atomicAdd(&S.cursor[cloud][(size_t)b * S.T[cloud] + index_bucket(S, cloud, 0, p)], 1);
        atomicAdd(&S.cursor[cloud + 2][(size_t)b * S.T[cloud + 2] + index_bucket(S, cloud, 1, p)], 1);
        // is the cloud ring-monotone? (LO:504-553 assumes it; the ring-azimuth search needs it)
        const int r0 = (int)p.w, r1 = i + 1 < n ? (int)pts[i + 1].w : r0;
        if (r0 < 0 || r0 >= S.rings || r1 < r0) { if (cloud == 0) L.mono_corner = 0; else L.mono_surf = 0; }
    }
}
__global__ void __launch_bounds__(256) k_index_partial(IndexSet S)
{
    __shared__ int ws[40];
    const int b = blockIdx.y;
    int t = 0;
    while (t < 3 && (int)blockIdx.x >= S.chunk_begin[t + 1]) ++t;
    const int chunk = blockIdx.x - S.chunk_begin[t], nchunk = S.T[t] / GRID_CHUNK;
    const int* cur = S.cursor[t] + (size_t)b * S.T[t] + (size_t)chunk * GRID_CHUNK;
    int s = 0;
#pragma unroll
    for (int k = 0; k < GRID_CHUNK / 256; ++k) s += cur[threadIdx.x * (GRID_CHUNK / 256) + k];
    int tot = 0;
    block_exclusive_scan(s, ws, &tot);
    if (threadIdx.x == 0) S.partial[t][(size_t)b * (nchunk + 1) + chunk] = tot;
}
__global__ void __launch_bounds__(256) k_index_scan(IndexSet S)
{
    __shared__ int ws[40];
    __shared__ int base_s;
    const int b = blockIdx.y;
    int t = 0;
    while (t < 3 && (int)blockIdx.x >= S.chunk_begin[t + 1]) ++t;
    const int chunk = blockIdx.x - S.chunk_begin[t], nchunk = S.T[t] / GRID_CHUNK, T = S.T[t];
    if (threadIdx.x < 32) {
        int v = 0;
        for (int q = threadIdx.x; q < chunk; q += 32) v += S.partial[t][(size_t)b * (nchunk + 1) + q];
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(LL_FULL_MASK, v, d);
        if (threadIdx.x == 0) base_s = v;
    }
    int* cur = S.cursor[t] + (size_t)b * T + (size_t)chunk * GRID_CHUNK;
    int* st = S.start[t] + (size_t)b * (T + 1) + (size_t)chunk * GRID_CHUNK;
    const int per = GRID_CHUNK / 256, i0 = threadIdx.x * per;
    int s = 0;
#pragma unroll
    for (int k = 0; k < per; ++k) s += cur[i0 + k];
    int tot = 0;
    int run = block_exclusive_scan(s, ws, &tot) + base_s;
#pragma unroll
    for (int k = 0; k < per; ++k) { const int c = cur[i0 + k]; st[i0 + k] = run; cur[i0 + k] = run; run += c; }
    if (chunk == nchunk - 1 && threadIdx.x == 255) S.start[t][(size_t)b * (T + 1) + T] = run;
}
__global__ void k_index_scatter(IndexSet S, const LaneState* lane)
{
    const int b = blockIdx.y, cloud = blockIdx.z;
    const LaneState& L = lane[b];
    const int n = cloud == 0 ? L.n_less_sharp : L.n_less_flat;
    const float4* pts = S.pts[cloud][L.cur] + (size_t)b * S.lane_stride[cloud];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = pts[i];
        int ring = (int)p.w;
        ring = ring < 0 ? 0 : (ring > 255 ? 255 : ring);
        const float4 rec = make_float4(p.x, p.y, p.z, __int_as_float((i & 0xFFFFFF) | (ring << 24)));
#pragma unroll
        for (int kind = 0; kind < 2; ++kind) {
            const int t = cloud + 2 * kind;
            const int pos = atomicAdd(&S.cursor[t][(size_t)b * S.T[t] + index_bucket(S, cloud, kind, p)], 1);
            S.sorted[t][(size_t)b * S.cap[t] + pos] = rec;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// association
// ------------------------------------------------------------------------------------------------------
struct OdomParams {
    LaneState* lane;
    const float4* sharp;   // [B][R*12]
    const float4* flat;    // [B][R*24]
    const float4* lsharp[2];
    const float4* lflat[2];
    int Nmax, R;
    KnnGrid gc, gs;
    KnnGrid ac, as_;       // ring x azimuth-bin indexes
    int az_bins_corner, az_bins_surf;
    int* corner_assoc;     // [B][R*12][2]
    int* plane_assoc;      // [B][R*24][4]
    double* blocks;        // [B][LL_BLOCK_DOUBLES][nblk_cap]
    int nblk_cap;
    int graph_from_frame;
    float vote_t_min;
    int outer;             // opti_counter
    int dev_skip;          // development only: bit mask of association stages to skip (timing experiments)
    int plane_shells;      // grid shells tried for the 2nd / 3rd plane neighbour before the literal walk
};

__device__ __forceinline__ int last_slot(const LaneState& L) { return L.last_slot; }  // previous frame's clouds

__global__ void __launch_bounds__(256, 6) k_odom_assoc(OdomParams P)
{
    const int b = blockIdx.y;
    const LaneState& L = P.lane[b];
    if (!L.inited) return;
    const int lane = lane_id();
    const int q = blockIdx.x * (blockDim.x >> 5) + warp_id();
    const int ns = L.n_sharp, nf = L.n_flat;
    if (q >= ns + nf) return;
    const bool is_corner = q < ns;
    if ((P.dev_skip & 2) && is_corner) return;
    if ((P.dev_skip & 4) && !is_corner) return;
    const int i = is_corner ? q : q - ns;
    const float4 p = is_corner ? P.sharp[(size_t)b * P.R * LL_SHARP_PER_RING + i] : P.flat[(size_t)b * P.R * LL_FLAT_PER_RING + i];

    // TransformToStart, LO:77-95 with DISTORTION 0: slerp(1, q) = +-q, same rotation bit for bit
    double sx, sy, sz;
    quat_rotate(L.para_q, (double)p.x, (double)p.y, (double)p.z, sx, sy, sz);
    const float qx = (float)(sx + L.para_t[0]), qy = (float)(sy + L.para_t[1]), qz = (float)(sz + L.para_t[2]);

    const int slot = last_slot(L);
    const KnnGrid& G = is_corner ? P.gc : P.gs;
    GridView gv;
    gv.start = G.start + (size_t)b * (G.T + 1);
    gv.sorted = G.sorted + (size_t)b * G.cap;
    gv.Tmask = G.T - 1;
    gv.h = G.h;
    gv.inv_h = G.inv_h;
    const float4* last = is_corner ? P.lsharp[slot] + (size_t)b * P.R * LL_LSHARP_PER_RING : P.lflat[slot] + (size_t)b * P.Nmax;
    const int n = is_corner ? L.n_last_corner : L.n_last_surf;

    // ---- 1-NN (kdtree*Last->nearestKSearch(pointSel, 1, ...), LO:494 / LO:656) -----------------------------------
    const float geps = 1e-3f;
    const int cx = (int)floorf(qx * gv.inv_h), cy = (int)floorf(qy * gv.inv_h), cz = (int)floorf(qz * gv.inv_h);
    const int smax = (int)ceilf((5.0f + geps) * gv.inv_h);
    u64 best[1];
    best[0] = ~0ull;
    if (n > 0) {
        u64 lb = ~0ull;  // lane-local best (d2 bits << 32 | index)
        auto upd = [&](const float4 t) {
            const float d2 = sqdist3(qx, qy, qz, t.x, t.y, t.z);
            const u64 key = ((u64)__float_as_uint(d2) << 32) | ((unsigned)__float_as_int(t.w) & 0xFFFFFFu);
            if (key < lb) lb = key;
        };
        grid_visit_near_pruned(gv, qx, qy, qz, cx, cy, cz, upd, [&]() {
            const u64 m = warp_min_u64(lb);
            return m == ~0ull ? INFINITY : __uint_as_float((unsigned)(m >> 32));
        });
        for (int s = 2;; ++s) {
            const u64 m = warp_min_u64(lb);
            const float safe = (float)(s - 1) * gv.h - geps;
            if ((m != ~0ull && __uint_as_float((unsigned)(m >> 32)) < safe * safe) || s > smax) { best[0] = m; break; }
            grid_visit_shell(gv, cx, cy, cz, s, upd);
        }
    }
    int closest = -1, ind2 = -1, ind3 = -1;
    if (P.dev_skip & 1) best[0] = ~0ull;
    if (best[0] != ~0ull && (double)__uint_as_float((unsigned)(best[0] >> 32)) < 25.0) {  // LO:497 / LO:659
        closest = (int)(unsigned)best[0];
        const int cring = (int)last[closest].w;  // int(intensity), LO:500 / LO:664
        u64 k2 = ~0ull, k3 = ~0ull;              // (d2 bits << 32) | visit order  -> strict '<' of the serial loops
        const unsigned down_base = (unsigned)n;
        bool resolved = false;
        if ((is_corner ? L.mono_corner : L.mono_surf) && !(P.dev_skip & 8)) {
            // Ring-monotone cloud (always the case for clouds produced by scanRegistration): the serial loops of
            // LO:504-553 / LO:668-721 visit exactly the points whose ring lies in [cring-2, cring+2].  Those rings
            // are searched through the ring x azimuth-bin index: a point whose azimuth differs from the query's by
            // D lies at least rho * sin(D) away (rho = horizontal range of the query), so bins are visited outward
            // from the query's azimuth until that bound exceeds the best distances found (or 5 m, LO:29).
            // Ties keep the loops' visit order through the rank in the key.
            const KnnGrid& A = is_corner ? P.ac : P.as_;
            const int NB = is_corner ? P.az_bins_corner : P.az_bins_surf;
            GridView av;
            av.start = A.start + (size_t)b * (A.T + 1);
            av.sorted = A.sorted + (size_t)b * A.cap;
            av.Tmask = A.T - 1;
            av.h = 0.f; av.inv_h = 0.f;
            auto consider = [&](const float4 t) {
                const unsigned bits = (unsigned)__float_as_int(t.w);
                const int j = (int)(bits & 0xFFFFFFu), rj = (int)(bits >> 24);
                const float d2 = sqdist3(t.x, t.y, t.z, qx, qy, qz);
                if (!(d2 < 25.0f) || j == closest) return;  // 25 is exact in fp32: same decision as the fp64 compare of LO:521
                const unsigned rank = j > closest ? (unsigned)(j - (closest + 1)) : down_base + (unsigned)(closest - 1 - j);
                const u64 key = ((u64)__float_as_uint(d2) << 32) | rank;
                if (rj == cring) { if (!is_corner && key < k2) k2 = key; }
                else if (is_corner) { if (key < k2) k2 = key; }
                else if (key < k3) k3 = key;
            };
            const float rho = sqrtf(qx * qx + qy * qy);
            const int b0 = azimuth_bin(qx, qy, NB);
            const float wbin = 6.2831853f / (float)NB;
            // lane -> (ring slot 0..4 = cring-2..cring+2, bin slot 0..5); 30 lanes per round
            const int rslot = lane / 6, bslot = lane % 6;
            const int ring = cring - 2 + rslot;
            const bool ring_ok = lane < 30 && ring >= 0 && ring < P.R && !(is_corner && ring == cring);
            int done = -1;  // offsets |k| <= done have been visited
            for (int round = 0;; ++round) {
                // round 0: offsets -2..+2 (bslot 0..4); round r >= 1: offsets +-(3r .. 3r+2)
                int off = 0;
                bool use = ring_ok;
                if (round == 0) { use = use && bslot < 5; off = bslot - 2; }
                else { const int mag = 3 * round + (bslot % 3); off = bslot < 3 ? mag : -mag; }
                const int reach = round == 0 ? 2 : 3 * round + 2;
                if (2 * reach + 1 > NB) {  // wrapped all the way round: keep each bin once
                    if (abs(off) > NB / 2 || (off == -(NB / 2))) use = false;
                }
                int bucket = -1 - lane;
                if (use && abs(off) <= NB / 2) bucket = ring * NB + ((b0 + off) % NB + NB) % NB;
                const unsigned grp = __match_any_sync(LL_FULL_MASK, bucket);
                if ((__ffs(grp) - 1) != lane) bucket = -1;
                grid_stream_buckets(av, bucket, consider);
                done = reach;
                const u64 m2 = warp_min_u64(k2), m3 = is_corner ? 0ull : warp_min_u64(k3);
                // every unvisited bin is at least done * wbin away in azimuth
                const float dmin = (float)done * wbin - 1e-4f;
                const float lb = (dmin >= 1.5707963f ? rho : rho * sinf(fmaxf(dmin, 0.f))) - 1e-3f;
                const float lb2 = lb > 0.f ? lb * lb : 0.f;
                const bool ok2 = m2 != ~0ull && __uint_as_float((unsigned)(m2 >> 32)) < lb2;
                const bool ok3 = is_corner || (m3 != ~0ull && __uint_as_float((unsigned)(m3 >> 32)) < lb2);
                if ((ok2 && ok3) || lb2 >= 25.0f || 2 * done + 1 >= NB) break;
            }
            resolved = true;
            if (lane == 0 && P.outer == 2) atomicAdd(&P.lane[b].dbg[0], 1);
        }
        if (!resolved) {
        // increasing scan line (LO:504-527 / LO:668-693); lane order inside a chunk = visit order.
        // Four chunks are fetched per iteration (loads first, then consumed in visit order until the first break).
        {
            bool stop = false;
            for (int j0 = closest + 1; j0 < n && !stop; j0 += 128) {
                float4 tv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int j = j0 + u * 32 + lane; if (j < n) tv[u] = last[j]; }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (stop || j0 + u * 32 >= n) continue;  // warp-uniform
                    const int j = j0 + u * 32 + lane;
                    bool brk = false;
                    u64 c2 = ~0ull, c3 = ~0ull;
                    if (j < n) {
                        const float4 t = tv[u];
                        const int rj = (int)t.w;
                        brk = rj > cring + 2;   // int ring vs cring + 2.5 (LO:511)
                        const float d2 = sqdist3(t.x, t.y, t.z, qx, qy, qz);
                        const u64 key = ((u64)__float_as_uint(d2) << 32) | (unsigned)(j - (closest + 1));
                        if (!brk && d2 < 25.0f) {
                            if (is_corner) {
                                if (!(rj <= cring)) c2 = key;  // LO:507: same scan line -> continue
                            } else {
                                if (rj <= cring) c2 = key;     // LO:682
                                else c3 = key;                 // LO:688
                            }
                        }
                    }
                    const unsigned bm = __ballot_sync(LL_FULL_MASK, brk);
                    const int fb = bm ? __ffs(bm) - 1 : 32;
                    if (lane < fb) { if (c2 < k2) k2 = c2; if (c3 < k3) k3 = c3; }
                    if (bm) stop = true;
                }
            }
        }
        // decreasing scan line (LO:530-553 / LO:696-721); visit order continues after the up-scan
        {
            bool stop = false;
            for (int j0 = closest - 1; j0 >= 0 && !stop; j0 -= 128) {
                float4 tv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { const int j = j0 - u * 32 - lane; if (j >= 0) tv[u] = last[j]; }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (stop || j0 - u * 32 < 0) continue;
                    const int j = j0 - u * 32 - lane;
                    bool brk = false;
                    u64 c2 = ~0ull, c3 = ~0ull;
                    if (j >= 0) {
                        const float4 t = tv[u];
                        const int rj = (int)t.w;
                        brk = rj < cring - 2;   // LO:537
                        const float d2 = sqdist3(t.x, t.y, t.z, qx, qy, qz);
                        const u64 key = ((u64)__float_as_uint(d2) << 32) | (down_base + (unsigned)(closest - 1 - j));
                        if (!brk && d2 < 25.0f) {
                            if (is_corner) {
                                if (!(rj >= cring)) c2 = key;
                            } else {
                                if (rj >= cring) c2 = key;
                                else c3 = key;
                            }
                        }
                    }
                    const unsigned bm = __ballot_sync(LL_FULL_MASK, brk);
                    const int fb = bm ? __ffs(bm) - 1 : 32;
                    if (lane < fb) { if (c2 < k2) k2 = c2; if (c3 < k3) k3 = c3; }
                    if (bm) stop = true;
                }
            }
        }
        }  // literal scan loops
        k2 = warp_min_u64(k2);
        k3 = warp_min_u64(k3);
        auto decode = [&](u64 k) -> int {
            if (k == ~0ull) return -1;
            const unsigned rk = (unsigned)k;
            return rk >= down_base ? closest - 1 - (int)(rk - down_base) : closest + 1 + (int)rk;
        };
        ind2 = decode(k2);
        ind3 = decode(k3);
    }
    if (lane == 0) {
        if (is_corner) {
            int* o = P.corner_assoc + ((size_t)b * P.R * LL_SHARP_PER_RING + i) * 2;
            o[0] = ind2 >= 0 ? closest : -1;  // LO:556
            o[1] = ind2;
        } else {
            int* o = P.plane_assoc + ((size_t)b * P.R * LL_FLAT_PER_RING + i) * 4;
            const bool ok = ind2 >= 0 && ind3 >= 0;  // LO:723
            o[0] = ok ? closest : -1;
            o[1] = ok ? ind2 : -1;
            o[2] = ok ? ind3 : -1;
            o[3] = 0;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// compaction + graph vote + residual-block records; one CTA per lane
// ------------------------------------------------------------------------------------------------------
#define PREP_THREADS 1024
__global__ void __launch_bounds__(PREP_THREADS) k_odom_prep(OdomParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // src / tgt xyz of the compacted plane matches, their feature index and votes
    const int maxp = P.R * LL_FLAT_PER_RING;
    float* srcx = reinterpret_cast<float*>(smem_raw);
    float* srcy = srcx + maxp;
    float* srcz = srcy + maxp;
    float* tgtx = srcz + maxp;
    float* tgty = tgtx + maxp;
    float* tgtz = tgty + maxp;
    int* fidx = reinterpret_cast<int*>(tgtz + maxp);
    float* wsel = reinterpret_cast<float*>(fidx + maxp);
    __shared__ int ws[40];

    const int b = blockIdx.x, tid = threadIdx.x;
    LaneState& L = P.lane[b];
    if (!L.inited) { if (tid == 0) L.n_blocks = 0; return; }
    const int ns = L.n_sharp, nf = L.n_flat, slot = last_slot(L);
    const float4* sharp = P.sharp + (size_t)b * P.R * LL_SHARP_PER_RING;
    const float4* flat = P.flat + (size_t)b * P.R * LL_FLAT_PER_RING;
    const float4* lastc = P.lsharp[slot] + (size_t)b * P.R * LL_LSHARP_PER_RING;
    const float4* lasts = P.lflat[slot] + (size_t)b * P.Nmax;
    const int* ca = P.corner_assoc + (size_t)b * P.R * LL_SHARP_PER_RING * 2;
    int* pa = P.plane_assoc + (size_t)b * P.R * LL_FLAT_PER_RING * 4;
    double* blk = P.blocks + (size_t)b * LL_BLOCK_DOUBLES * P.nblk_cap;
    const int cap = P.nblk_cap;

    // ---- corners: compaction in feature order, LidarEdgeFactor records (LO:556-618) -------------------
    int ncorner = 0;
    {
        const int i = tid;  // R*12 <= 768 <= PREP_THREADS
        const bool v = i < ns && ca[i * 2 + 1] >= 0;
        const int pos = block_exclusive_scan(v ? 1 : 0, ws, &ncorner);
        if (v) {
            const float4 cp = sharp[i], a = lastc[ca[i * 2]], c = lastc[ca[i * 2 + 1]];
            blk[0 * cap + pos] = 0.0;
            blk[1 * cap + pos] = cp.x; blk[2 * cap + pos] = cp.y; blk[3 * cap + pos] = cp.z;
            blk[4 * cap + pos] = a.x; blk[5 * cap + pos] = a.y; blk[6 * cap + pos] = a.z;
            blk[7 * cap + pos] = c.x; blk[8 * cap + pos] = c.y; blk[9 * cap + pos] = c.z;
            blk[10 * cap + pos] = 1.0;
        }
    }
    // ---- planes: compaction (R*24 <= 1536: up to 2 per thread, contiguous chunks keep the order) --------
    const int per = (maxp + PREP_THREADS - 1) / PREP_THREADS;
    const int i0 = tid * per, i1 = min(i0 + per, nf);
    int mine = 0;
    for (int i = i0; i < i1; ++i) mine += pa[i * 4] >= 0;
    int nplane = 0;
    int pos = block_exclusive_scan(mine, ws, &nplane);
    for (int i = i0; i < i1; ++i) {
        if (pa[i * 4] >= 0) {
            const float4 s = flat[i], t = lasts[pa[i * 4]];  // Corre_Match src / tgt, LO:753-754
            srcx[pos] = s.x; srcy[pos] = s.y; srcz[pos] = s.z;
            tgtx[pos] = t.x; tgty[pos] = t.y; tgtz[pos] = t.z;
            fidx[pos] = i;
            ++pos;
        }
    }
    __syncthreads();
    // ---- graph_based_correspondence_vote_simple, plane case: 10 contiguous regions (LO:184-252) ----------
    const bool vote = L.now_frame > P.graph_from_frame;  // LO:781 / LO:794
    const int region_len = nplane / 10;
    for (int k = tid; k < nplane; k += PREP_THREADS) {
        float w = 1.0f;  // LO:783: weight 1 while now_frame <= 5
        if (vote) {
            int reg = region_len > 0 ? k / region_len : 9;
            if (reg > 9) reg = 9;
            const int r0 = region_len * reg, r1 = reg == 9 ? nplane : region_len * (reg + 1);
            const float ax = srcx[k], ay = srcy[k], az = srcz[k], bx = tgtx[k], by = tgty[k], bz = tgtz[k];
            int votes = 0;
            for (int j = r0; j < r1; ++j) {
                if (j == k) continue;
                // Distance() LO:153-162 is symmetric bit for bit, so each unordered pair is evaluated from both ends
                const float s1 = sqrtf(sqdist3(ax, ay, az, srcx[j], srcy[j], srcz[j]));
                const float s2 = sqrtf(sqdist3(bx, by, bz, tgtx[j], tgty[j], tgtz[j]));
                const float gap = fabsf(s1 - s2);
                // score = expf(-(gap*gap)/(1*1)) < 0.96f  <=>  gap*gap >= t_min (host-calibrated on glibc expf)
                votes += (gap * gap >= P.vote_t_min);
            }
            const float num_selected = 0.90f * (float)(r1 - r0);  // LO:299-300
            if ((float)votes > num_selected) w = 0.f;             // LO:312-316: this and all worse are dropped
            else if ((float)votes <= 50.f) w = 5.0f;              // LO:317-318
            else w = 1.0f;
        }
        wsel[k] = w;
        pa[fidx[k] * 4 + 3] = (int)(w * 1000.f);
    }
    __syncthreads();
    // ---- LidarPlaneFactor_modify records for the selected matches (LO:797-808 / LO:781-787) ---------------
    const int k0 = tid * per, k1 = min(k0 + per, nplane);
    int sel = 0;
    for (int k = k0; k < k1; ++k) sel += wsel[k] > 0.f;
    int nsel = 0;
    int o = ncorner + block_exclusive_scan(sel, ws, &nsel);
    for (int k = k0; k < k1; ++k) {
        if (wsel[k] > 0.f) {
            const int i = fidx[k];
            const float4 cp = flat[i], pj = lasts[pa[i * 4]], pl = lasts[pa[i * 4 + 1]], pm = lasts[pa[i * 4 + 2]];
            // LF:210-211  ljm_norm = (j - l).cross(j - m); normalize()
            const double ax = (double)pj.x - (double)pl.x, ay = (double)pj.y - (double)pl.y, az = (double)pj.z - (double)pl.z;
            const double bx = (double)pj.x - (double)pm.x, by = (double)pj.y - (double)pm.y, bz = (double)pj.z - (double)pm.z;
            double nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
            const double z = nx * nx + ny * ny + nz * nz;
            if (z > 0.0) { const double nn = sqrt(z); nx = nx / nn; ny = ny / nn; nz = nz / nn; }
            blk[0 * cap + o] = 1.0;
            blk[1 * cap + o] = cp.x; blk[2 * cap + o] = cp.y; blk[3 * cap + o] = cp.z;
            blk[4 * cap + o] = pj.x; blk[5 * cap + o] = pj.y; blk[6 * cap + o] = pj.z;
            blk[7 * cap + o] = nx; blk[8 * cap + o] = ny; blk[9 * cap + o] = nz;
            blk[10 * cap + o] = (double)wsel[k];
            ++o;
        }
    }
    if (tid == 0) {
        L.n_blocks = ncorner + nsel;
        L.n_corner_corr = ncorner;
        L.n_plane_corr = nplane;
        L.n_plane_sel = nsel;
        L.corner_corr[P.outer] = ncorner;
        L.plane_corr[P.outer] = nplane;
        L.plane_sel[P.outer] = nsel;
    }
}

__global__ void __launch_bounds__(LM_THREADS) k_lm_solve_odom(OdomParams P)
{
    const int b = blockIdx.x;
    LaneState& L = P.lane[b];
    if (!L.inited) return;
    lm_solve(P.blocks + (size_t)b * LL_BLOCK_DOUBLES * P.nblk_cap, P.nblk_cap, L.n_blocks, L.para_q, L.para_t, &L, P.outer);
}

// LO:830-831 pose accumulation; LO:882-896 swap (the grids are rebuilt right after); counters LO:925-926
__global__ void k_odom_finalize(LaneState* lane, double* pose_out, int n_lanes)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_lanes) return;
    LaneState& L = lane[b];
    if (!L.inited) {
        L.inited = 1;  // LO:427-431
    } else {
        double rx, ry, rz;
        quat_rotate(L.q_w, L.para_t[0], L.para_t[1], L.para_t[2], rx, ry, rz);
        L.t_w[0] = L.t_w[0] + rx;
        L.t_w[1] = L.t_w[1] + ry;
        L.t_w[2] = L.t_w[2] + rz;
        double qn[4];
        quat_mul(L.q_w, L.para_q, qn);
        L.q_w[0] = qn[0]; L.q_w[1] = qn[1]; L.q_w[2] = qn[2]; L.q_w[3] = qn[3];
    }
    L.last_slot = L.cur;  // LO:882-891 swap
    L.mono_corner = 1;    // cleared by k_grid_count if the new *Last clouds are not ring-monotone
    L.mono_surf = 1;
    L.n_last_corner = L.n_less_sharp;
    L.n_last_surf = L.n_less_flat;
    L.now_frame++;
    if (pose_out) {
        double* o = pose_out + (size_t)b * 14;
        for (int k = 0; k < 4; ++k) { o[k] = L.q_w[k]; o[7 + k] = L.q_w[k]; }
        for (int k = 0; k < 3; ++k) { o[4 + k] = L.t_w[k]; o[11 + k] = L.t_w[k]; }
    }
}

}  // namespace

static int build_grid(ll_ctx* c, KnnGrid& g, const float4* p0, const float4* p1, size_t lane_stride, int which, int n_lanes, int max_pts, int az_bins = 0)
{
    GridSrc S;
    S.pts[0] = p0; S.pts[1] = p1; S.lane_stride = lane_stride; S.which = which; S.az_bins = az_bins; S.rings = c->R;
    cudaStream_t s = c->stream;
    LL_CUDA_CHECK(c, cudaMemsetAsync(g.cursor, 0, sizeof(int) * (size_t)g.T * n_lanes, s));
    const int gx = (max_pts + 255) / 256 > 0 ? (max_pts + 255) / 256 : 1;
    { LLProf pr(c, "k_grid_count"); k_grid_count<<<dim3(gx < 296 ? gx : 296, n_lanes), 256, 0, s>>>(S, c->d_lane, g.cursor, g.T, g.inv_h); }
    { LLProf pr(c, "k_grid_partial"); k_grid_partial<<<dim3(g.T / GRID_CHUNK, n_lanes), 256, 0, s>>>(g.cursor, g.partial, g.T); }
    { LLProf pr(c, "k_grid_scan"); k_grid_scan<<<dim3(g.T / GRID_CHUNK, n_lanes), 256, 0, s>>>(g.cursor, g.start, g.partial, g.T); }
    { LLProf pr(c, "k_grid_scatter"); k_grid_scatter<<<dim3(gx < 296 ? gx : 296, n_lanes), 256, 0, s>>>(S, c->d_lane, g.cursor, g.sorted, g.T, g.cap, g.inv_h); }
    c->launches += 4;
    return LL_OK;
}

int ll_launch_odometry(ll_ctx* c, int n_lanes)
{
    OdomParams P;
    P.lane = c->d_lane; P.sharp = c->d_sharp; P.flat = c->d_flat;
    P.lsharp[0] = c->d_lsharp[0]; P.lsharp[1] = c->d_lsharp[1]; P.lflat[0] = c->d_lflat[0]; P.lflat[1] = c->d_lflat[1];
    P.Nmax = c->Nmax; P.R = c->R; P.gc = c->g_corner; P.gs = c->g_surf; P.ac = c->a_corner; P.as_ = c->a_surf; P.az_bins_corner = c->az_bins_corner; P.az_bins_surf = c->az_bins_surf;
    P.corner_assoc = c->d_corner_assoc; P.plane_assoc = c->d_plane_assoc; P.blocks = c->d_blocks; P.nblk_cap = c->nblk_cap;
    P.graph_from_frame = c->cfg.graph_from_frame; P.vote_t_min = c->vote_t_min; P.plane_shells = c->plane_shells; P.dev_skip = getenv("LL_DEV_SKIP") ? atoi(getenv("LL_DEV_SKIP")) : 0;
    cudaStream_t s = c->stream;
    const int nq = c->R * (LL_SHARP_PER_RING + LL_FLAT_PER_RING);
    const size_t prep_smem = (size_t)c->R * LL_FLAT_PER_RING * 8 * 4;
    LL_CUDA_CHECK(c, cudaFuncSetAttribute(k_odom_prep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prep_smem));
    for (int outer = 0; outer < 3; ++outer) {  // LO:439
        P.outer = outer;
        { LLProf pr(c, "k_odom_assoc"); k_odom_assoc<<<dim3((nq + 7) / 8, n_lanes), 256, 0, s>>>(P); }
        { LLProf pr(c, "k_odom_prep"); k_odom_prep<<<n_lanes, PREP_THREADS, prep_smem, s>>>(P); }
        { LLProf pr(c, "k_lm_solve_odom"); k_lm_solve_odom<<<n_lanes, LM_THREADS, 0, s>>>(P); }
        c->launches += 3;
    }
    { LLProf pr(c, "k_odom_finalize"); k_odom_finalize<<<(n_lanes + 63) / 64, 64, 0, s>>>(c->d_lane, c->d_pose, n_lanes); }
    c->launches += 1;
    // kdtreeCornerLast / kdtreeSurfLast ->setInputCloud on this frame's less-sharp / less-flat (LO:895-896):
    // the spatial hash grids and the ring x azimuth indexes of both clouds are built together, 4 launches
    {
        IndexSet S;
        KnnGrid* tab[4] = {&c->g_corner, &c->g_surf, &c->a_corner, &c->a_surf};
        S.pts[0][0] = c->d_lsharp[0]; S.pts[0][1] = c->d_lsharp[1]; S.pts[1][0] = c->d_lflat[0]; S.pts[1][1] = c->d_lflat[1];
        S.lane_stride[0] = (size_t)c->R * LL_LSHARP_PER_RING; S.lane_stride[1] = (size_t)c->Nmax;
        S.chunk_begin[0] = 0;
        for (int t = 0; t < 4; ++t) {
            S.cursor[t] = tab[t]->cursor; S.start[t] = tab[t]->start; S.partial[t] = tab[t]->partial; S.sorted[t] = tab[t]->sorted;
            S.T[t] = tab[t]->T; S.cap[t] = tab[t]->cap;
            S.chunk_begin[t + 1] = S.chunk_begin[t] + tab[t]->T / GRID_CHUNK;
            LL_CUDA_CHECK(c, cudaMemsetAsync(tab[t]->cursor, 0, sizeof(int) * (size_t)tab[t]->T * n_lanes, s));
        }
        S.inv_h = c->g_corner.inv_h; S.az_bins[0] = c->az_bins_corner; S.az_bins[1] = c->az_bins_surf; S.rings = c->R;
        const int gx = (c->Nmax / 2 + 255) / 256 < 148 ? (c->Nmax / 2 + 255) / 256 : 148;
        { LLProf pr(c, "k_index_count"); k_index_count<<<dim3(gx, n_lanes, 2), 256, 0, s>>>(S, c->d_lane); }
        { LLProf pr(c, "k_index_partial"); k_index_partial<<<dim3(S.chunk_begin[4], n_lanes), 256, 0, s>>>(S); }
        { LLProf pr(c, "k_index_scan"); k_index_scan<<<dim3(S.chunk_begin[4], n_lanes), 256, 0, s>>>(S); }
        { LLProf pr(c, "k_index_scatter"); k_index_scatter<<<dim3(gx, n_lanes, 2), 256, 0, s>>>(S, c->d_lane); }
        c->launches += 4;
    }
    LL_CUDA_CHECK(c, cudaGetLastError());
    return LL_OK;
}

// used by the mapping module: grids over the gathered local map (slot 0 of the given buffers)
int ll_build_map_grid(ll_ctx* c, KnnGrid& g, const float4* pts, size_t lane_stride, int which, int n_lanes, int max_pts)
{
    return build_grid(c, g, pts, pts, lane_stride, which, n_lanes, max_pts);
}
