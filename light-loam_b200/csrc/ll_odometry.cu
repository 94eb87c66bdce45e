// Scan-to-scan odometry on the device — replaces laserOdometry.cpp:425-896.
//
//   k_index_*        kdtree*Last->setInputCloud (LO:895-896) -> polar index (azimuth bin x ring) + ring elevation bands
//                    (k_grid_* below build the hashed grids of the mapping module, LM:1830-1831)
//   k_odom_assoc     LO:491-556 + LO:653-723: TransformToStart, exact 1-NN (d2 < 25), ring-window 2nd / 3rd point;
//                    one thread per feature point, k_odom_assoc_heavy: one warp per query that needs a wide search
//   k_odom_prep      order-preserving compaction of the matches, residual-block records (LF ctor maths)
//   k_odom_vote      graph_based_correspondence_vote_simple (LO:165-342) for planes when now_frame > 5, one CTA per region
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
};
__device__ __forceinline__ int src_bucket(const GridSrc& S, const float4 p, int T, float inv_h)
{
    return cell_bucket((int)floorf(p.x * inv_h), (int)floorf(p.y * inv_h), (int)floorf(p.z * inv_h), T - 1);
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

// ---- fused build of the two odometry indexes (corner / surf: ring x azimuth-bin tables) ---------------------------
// The index is polar, in the frame the *Last cloud was measured in: bucket = azimuth bin * R + ring (bin-major, so
// the rings of one bin are consecutive buckets and any ring range of a bin is ONE contiguous span of the sorted
// array).  Next to the buckets the build records, per ring, the smallest and largest elevation angle of the ring's
// points; a query at elevation e is at least |q| sin(gap) away from every point of a ring whose elevation band lies
// `gap` away from e, and at least rho sin(D) away from every point whose azimuth differs by D — the two bounds that
// make the search exact (k_odom_assoc).
struct IndexSet {
    const float4* pts[2][2];   // [cloud][ping-pong slot]
    size_t lane_stride[2];
    int* cursor[2];            // table = cloud (0 corner, 1 surf)
    int* start[2];
    int* partial[2];
    float4* sorted[2];
    unsigned* ebound[2];       // [B][R][2] order-preserving encodings: max of ~enc(tan elevation), max of enc(tan elevation); 0 = empty
    float* bands[2];           // [B][BAND_STRIDE] decoded: elo[R], ehi[R] (empty rings filled in), then the "bands are ordered" flag
    int T[2], cap[2];
    int chunk_begin[3];        // prefix of T / GRID_CHUNK over the tables
    int az_bins[2];
    int rings;
    int gx_corner;             // k_index_count / k_index_scatter: blocks [0, gx_corner) of a lane take the corner cloud, the rest the surf cloud
};
#define BAND_STRIDE (2 * LL_MAX_RINGS + 4)
__device__ __forceinline__ unsigned enc_f32(float f) { const unsigned b = __float_as_uint(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
__device__ __forceinline__ float dec_f32(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u); }
// tangent of the elevation angle: monotone in the elevation, so a ring's band can be tracked as min / max of this
// (one rsqrt per point) and converted to angles once per ring (k_index_partial)
__device__ __forceinline__ float elevation_tan(float x, float y, float z)
{
    const float r2 = x * x + y * y;
    if (!(r2 > 0.f)) return z > 0.f ? INFINITY : (z < 0.f ? -INFINITY : 0.f);
    return z * rsqrtf(r2);
}
__device__ __forceinline__ int index_ring(const IndexSet& S, const float4 p)
{
    int r = (int)p.w;
    return r < 0 ? 0 : (r >= S.rings ? S.rings - 1 : r);
}
__device__ __forceinline__ int index_bucket(const IndexSet& S, int cloud, const float4 p)
{
    return azimuth_bin(p.x, p.y, S.az_bins[cloud]) * S.rings + index_ring(S, p);
}
#define IDX_UNROLL 4
__global__ void k_index_count(IndexSet S, LaneState* lane)
{
    // one grid for both clouds, sized for what they usually hold (a block that finds nothing to do still pays the
    // launch and the lane-state load): the first gx_corner blocks of a lane take the corner cloud
    const int b = blockIdx.y, cloud = (int)blockIdx.x < S.gx_corner ? 0 : 1;
    const int bx = cloud == 0 ? blockIdx.x : blockIdx.x - S.gx_corner, gxc = cloud == 0 ? S.gx_corner : (int)gridDim.x - S.gx_corner;
    LaneState& L = lane[b];
    const int n = (!L.inited || L.err) ? 0 : (cloud == 0 ? L.n_last_corner : L.n_last_surf);   // laserCloudCornerLast / SurfLast
    const float4* pts = S.pts[cloud][L.last_slot] + (size_t)b * S.lane_stride[cloud];
    unsigned* eb = S.ebound[cloud] + (size_t)b * S.rings * 2;
    // a warp takes IDX_UNROLL chunks of 32 consecutive points per round; their loads are all issued before the first is used
    const int ln = threadIdx.x & 31;
    for (int base0 = (bx * blockDim.x + (threadIdx.x & ~31)) * IDX_UNROLL; base0 < n; base0 += gxc * blockDim.x * IDX_UNROLL) {
        float4 pv[IDX_UNROLL];
        float wn[IDX_UNROLL];
#pragma unroll
        for (int u = 0; u < IDX_UNROLL; ++u) {
            const int i = base0 + u * 32 + ln;
            if (i < n) { pv[u] = pts[i]; wn[u] = i + 1 < n ? pts[i + 1].w : pv[u].w; }
        }
#pragma unroll
        for (int u = 0; u < IDX_UNROLL; ++u) {
            const int i = base0 + u * 32 + ln;
            if (base0 + u * 32 >= n) break;   // warp-uniform
            int ring = -1, bk = -1 - ln;      // distinct invalid ids: match_any never groups idle lanes
            unsigned elo = 0u, ehi = 0u;
            if (i < n) {
                const float4 p = pv[u];
                bk = index_bucket(S, cloud, p);
                // is the cloud ring-monotone? (LO:504-553 assumes it; the ring-window search needs it)
                const int r0 = (int)p.w, r1 = (int)wn[u];
                if (r0 < 0 || r0 >= S.rings || r1 < r0) { if (cloud == 0) L.mono_corner = 0; else L.mono_surf = 0; }
                ring = index_ring(S, p);
                const unsigned e = enc_f32(elevation_tan(p.x, p.y, p.z));
                elo = ~e; ehi = e;
            }
            // neighbouring points of the cloud mostly share their bucket: one atomic per distinct bucket of the warp
            const unsigned grp = __match_any_sync(LL_FULL_MASK, bk);
            if (bk >= 0 && ln == __ffs(grp) - 1) atomicAdd(&S.cursor[cloud][(size_t)b * S.T[cloud] + bk], __popc(grp));
            // a warp's 32 consecutive points nearly always share the ring: one pair of atomics per warp
            const unsigned have = __ballot_sync(LL_FULL_MASK, ring >= 0);
            if (!have) continue;
            const int ring0 = __shfl_sync(LL_FULL_MASK, ring, __ffs(have) - 1);
            if (__all_sync(LL_FULL_MASK, ring < 0 || ring == ring0)) {
                const unsigned mlo = __reduce_max_sync(LL_FULL_MASK, elo), mhi = __reduce_max_sync(LL_FULL_MASK, ehi);
                if (ln == 0) { atomicMax(&eb[ring0 * 2], mlo); atomicMax(&eb[ring0 * 2 + 1], mhi); }
            } else if (ring >= 0) {
                atomicMax(&eb[ring * 2], elo); atomicMax(&eb[ring * 2 + 1], ehi);
            }
        }
    }
}
__global__ void __launch_bounds__(256) k_index_partial(IndexSet S, LaneState* lane)
{
    __shared__ int ws[40];
    const int b = blockIdx.y;
    const int t = (int)blockIdx.x >= S.chunk_begin[1] ? 1 : 0;
    const int chunk = blockIdx.x - S.chunk_begin[t], nchunk = S.T[t] / GRID_CHUNK;
    const int* cur = S.cursor[t] + (size_t)b * S.T[t] + (size_t)chunk * GRID_CHUNK;
    int s = 0;
#pragma unroll
    for (int k = 0; k < GRID_CHUNK / 256; ++k) s += cur[threadIdx.x * (GRID_CHUNK / 256) + k];
    int tot = 0;
    block_exclusive_scan(s, ws, &tot);
    if (threadIdx.x == 0) S.partial[t][(size_t)b * (nchunk + 1) + chunk] = tot;
    // the first chunk of each table also decodes the ring elevation bands (complete since k_index_count): thread r turns
    // ring r's tangents into angles, then one thread walks the rings.  An empty ring gets an empty band at its
    // predecessor's upper edge, so ordered bands stay ordered through the gaps.
    if (chunk == 0 && t == 0 && threadIdx.x == 0) {   // once per lane and step: the selection counters the vote CTAs add to
        LaneState& L = lane[b];
        L.plane_sel[0] = L.plane_sel[1] = L.plane_sel[2] = 0;
    }
    if (chunk == 0) {   // block-uniform
        __shared__ float s_lo[LL_MAX_RINGS], s_hi[LL_MAX_RINGS];
        __shared__ unsigned char s_has[LL_MAX_RINGS];
        const unsigned* eb = S.ebound[t] + (size_t)b * S.rings * 2;
        float* bd = S.bands[t] + (size_t)b * BAND_STRIDE;
        const int r = threadIdx.x;
        if (r < S.rings) {
            const unsigned lo = eb[r * 2], hi = eb[r * 2 + 1];
            s_has[r] = hi != 0u;
            if (hi != 0u) { s_lo[r] = atanf(dec_f32(~lo)) - 1e-5f; s_hi[r] = atanf(dec_f32(hi)) + 1e-5f; }   // widened by the rsqrt / atanf error
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            float plo = -INFINITY, phi = -INFINITY;
            int ordered = 1;
            for (int q = 0; q < S.rings; ++q) {
                float elo = phi, ehi = phi;
                if (s_has[q]) {
                    elo = s_lo[q]; ehi = s_hi[q];
                    if (elo < plo || ehi < phi) ordered = 0;
                    plo = elo; phi = ehi;
                }
                s_lo[q] = elo; s_hi[q] = ehi;
            }
            reinterpret_cast<int*>(bd)[2 * LL_MAX_RINGS] = ordered;
        }
        __syncthreads();
        if (r < S.rings) { bd[r] = s_lo[r]; bd[LL_MAX_RINGS + r] = s_hi[r]; }
    }
}
__global__ void __launch_bounds__(256) k_index_scan(IndexSet S)
{
    __shared__ int ws[40];
    __shared__ int base_s;
    const int b = blockIdx.y;
    const int t = (int)blockIdx.x >= S.chunk_begin[1] ? 1 : 0;
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
    const int b = blockIdx.y, cloud = (int)blockIdx.x < S.gx_corner ? 0 : 1;
    const int bx = cloud == 0 ? blockIdx.x : blockIdx.x - S.gx_corner, gxc = cloud == 0 ? S.gx_corner : (int)gridDim.x - S.gx_corner;
    const LaneState& L = lane[b];
    const int n = (!L.inited || L.err) ? 0 : (cloud == 0 ? L.n_last_corner : L.n_last_surf);
    const float4* pts = S.pts[cloud][L.last_slot] + (size_t)b * S.lane_stride[cloud];
    const int ln = threadIdx.x & 31;
    for (int base0 = (bx * blockDim.x + (threadIdx.x & ~31)) * IDX_UNROLL; base0 < n; base0 += gxc * blockDim.x * IDX_UNROLL) {
        float4 pv[IDX_UNROLL];
#pragma unroll
        for (int u = 0; u < IDX_UNROLL; ++u) {
            const int i = base0 + u * 32 + ln;
            pv[u] = i < n ? pts[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < IDX_UNROLL; ++u) {
            const int i = base0 + u * 32 + ln;
            if (base0 + u * 32 >= n) break;   // warp-uniform
            const float4 p = pv[u];
            int bk = -1 - ln, ring = 0;
            if (i < n) { ring = index_ring(S, p); bk = index_bucket(S, cloud, p); }
            // one cursor atomic per distinct bucket of the warp; the order inside a bucket is free (every consumer takes
            // a minimum over a total order)
            const unsigned grp = __match_any_sync(LL_FULL_MASK, bk);
            const int leader = __ffs(grp) - 1;
            int pos = 0;
            if (bk >= 0 && ln == leader) pos = atomicAdd(&S.cursor[cloud][(size_t)b * S.T[cloud] + bk], __popc(grp));
            pos = __shfl_sync(LL_FULL_MASK, pos, leader) + __popc(grp & ((1u << ln) - 1u));
            if (bk >= 0) S.sorted[cloud][(size_t)b * S.cap[cloud] + pos] = make_float4(p.x, p.y, p.z, __int_as_float((i & 0xFFFFFF) | (ring << 24)));
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
    KnnGrid ac, as_;       // polar (azimuth bin x ring) indexes of the corner / surf *Last clouds
    const float* bands[2]; // [B][BAND_STRIDE] ring elevation bands of the two indexed clouds
    int az_bins_corner, az_bins_surf;
    int* corner_assoc;     // [B][R*12][2]
    int* plane_assoc;      // [B][R*24][4]
    double* blocks;        // [B][LL_BLOCK_DOUBLES][nblk_cap]
    int nblk_cap;
    int graph_from_frame;
    float vote_t_min;
    int distortion;        // LO:23 DISTORTION: 0 reference build; 1 per-point interpolation ratio; 2 additionally TransformToEnd (LO:861-880)
    int outer;             // opti_counter
    int dev_skip;          // development only: bit mask of association stages to skip (timing experiments)
    float4* vote_src;      // [B][R*24] compacted plane matches: current point (w = feature index) / closest point
    float4* vote_tgt;
    int4* queue;           // queries handed from the per-thread pass to the warp pass (k_odom_assoc_heavy)
    int* queue_n;          // per outer iteration: [0..2] long entries (stored from the front), [4..6] pop cursors, [8..10] short entries (from the back)
    int queue_cap;
};

__device__ __forceinline__ int last_slot(const LaneState& L) { return L.last_slot; }  // previous frame's clouds

// ------------------------------------------------------------------------------------------------------
// k_odom_assoc: one THREAD per feature point for the common case, the WARP for the rare long searches.
//
// Both searches of a query run over ONE polar index of the *Last cloud (bucket = azimuth bin * R + ring, built in the
// frame the cloud was measured in, so points are spread evenly over the buckets at every range):
//   * the exact nearest neighbour (kdtree*Last->nearestKSearch(pointSel, 1, ...), LO:494 / LO:656; accepted when
//     d2 < 25): a target whose direction is an angle g away from the query's lies at least |q| sin g away, and
//     g >= |elevation difference| as well as (for the horizontal projections) >= |azimuth difference|.  With the
//     per-ring elevation bands recorded by the index build this bounds, for the best distance found so far, the rings
//     and the azimuth bins that can still hold something closer; the search starts at the query's own bucket and stops
//     when nothing remains inside those bounds.
//   * the ring window of LO:504-553 / LO:668-721 (rings c-2 .. c+2 around the closest point's ring c).
// A query normally sees a few dozen candidates - far too few to feed a warp, so every lane runs its own query.
// Searches that do not resolve inside that budget (no target nearby, 2nd / 3rd point not bounded after a few bins,
// clouds that are not ring-sorted) are handed to the whole warp, one query at a time (k_odom_assoc_heavy).
// Decisions are the ones of LO:491-556 / LO:653-723: the minima are taken over total orders ((d2 bits, target
// index) for the 1-NN, (d2 bits, visit rank of the serial loops) for the 2nd / 3rd point), so the visiting order
// and the thread / warp split are free.
// ------------------------------------------------------------------------------------------------------
#define ASSOC_THREADS 128
// Read-only 16-byte load that asks L2 to bring in the whole 128-byte line (8 bucket-sorted points): the following
// candidates of the same bucket then hit L2 instead of paying another DRAM round trip.
__device__ __forceinline__ float4 ld_point(const float4* p)
{
    float4 v;
    asm volatile("ld.global.nc.L2::128B.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
// Where a search reads the polar index from.  StageGlobal: the lane's tables in global memory (bin = absolute azimuth
// bin).  StageShared: the part of the tables a CTA staged in shared memory (k_odom_assoc_slab): `nbins` consecutive
// bins starting at absolute bin `bin0`, headers rewritten to positions inside the staged point array.
struct StageGlobal {
    const int* __restrict__ start;
    const float4* __restrict__ sorted;
    int NB, R;
    __device__ __forceinline__ float4 point(int k) const { return ld_point(sorted + k); }
    __device__ __forceinline__ int header(int bin, int ring) const { return __ldg(start + (bin & (NB - 1)) * R + ring); }   // NB is a power of two
};
struct StageShared {
    unsigned hdr, pts;   // shared-window addresses
    int bin0, NB, R;
    __device__ __forceinline__ float4 point(int k) const
    {
        float4 v;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(pts + (unsigned)k * 16u));
        return v;
    }
    __device__ __forceinline__ int header(int bin, int ring) const
    {
        int v;
        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(hdr + (unsigned)((((bin - bin0) & (NB - 1)) * R + ring) * 4)));
        return v;
    }
};
// Up to three spans as ONE candidate stream (four loads in flight across span boundaries): the spans of neighbouring
// azimuth bins cost one memory round trip together instead of one each.
#ifndef ASSOC_W
#define ASSOC_W 4   // candidate loads in flight per thread and round trip
#endif
template <typename ST, typename F>
__device__ __forceinline__ void for_spans3(const ST& st, int s0, int e0, int s1, int e1, int s2, int e2, F&& f)
{
    const int n0 = e0 - s0, n01 = n0 + (e1 - s1), tot = n01 + (e2 - s2);
    const int d1 = s1 - n0, d2 = s2 - n01;   // stream position -> array index: + s0, + d1 or + d2 depending on the span
    auto at = [&](int k) { k = min(k, tot - 1); return k + (k < n0 ? s0 : (k < n01 ? d1 : d2)); };
#pragma unroll 1
    for (int k = 0; k < tot; k += ASSOC_W) {
        float4 t[ASSOC_W];
#pragma unroll
        for (int u = 0; u < ASSOC_W; ++u) t[u] = st.point(at(k + u));
#pragma unroll
        for (int u = 0; u < ASSOC_W; ++u) f(t[u]);
    }
}
// development statistics of the thread pass (compile with -DLL_ASSOC_STATS): candidates per phase and range class
#ifdef LL_ASSOC_STATS
__device__ unsigned long long g_assoc_stats[32];
#define ASTAT(slot, v) (stat[(slot)] += (v))
#else
#define ASTAT(slot, v) ((void)0)
#endif
struct AssocQuery {
    float qx, qy, qz;
    int closest, cring, n;
    u64 k2, k3;
    unsigned cut;   // d2 bit patterns above this cannot improve k2 / k3 (and 25.0f and above never count, LO:521)
};
#define D2_BITS_25 0x41C80000u   // 25.0f; for d2 >= 0 the bit patterns order like the values, NaNs sort above
template <bool CORNER>
__device__ __forceinline__ void assoc_update_cut(AssocQuery& Q)
{
    const unsigned h2 = (unsigned)(Q.k2 >> 32), h3 = (unsigned)(Q.k3 >> 32);
    Q.cut = min(D2_BITS_25 - 1u, CORNER ? h2 : max(h2, h3));
}
template <bool CORNER>
__device__ __forceinline__ void assoc_consider(AssocQuery& Q, const float4 t)
{
    const float d2 = sqdist3(t.x, t.y, t.z, Q.qx, Q.qy, Q.qz);
    if (__float_as_uint(d2) > Q.cut) return;   // most candidates: farther than what is already held (or >= 25, LO:521)
    const unsigned bits = (unsigned)__float_as_int(t.w);
    const int j = (int)(bits & 0xFFFFFFu), rj = (int)(bits >> 24);
    if (j == Q.closest) return;
    const unsigned rank = j > Q.closest ? (unsigned)(j - (Q.closest + 1)) : (unsigned)Q.n + (unsigned)(Q.closest - 1 - j);
    const u64 key = ((u64)__float_as_uint(d2) << 32) | rank;
    if (CORNER) { if (rj != Q.cring && key < Q.k2) { Q.k2 = key; assoc_update_cut<CORNER>(Q); } }  // same scan line -> continue (LO:507 / LO:533)
    else if (rj == Q.cring) { if (key < Q.k2) { Q.k2 = key; assoc_update_cut<CORNER>(Q); } }  // LO:682 / LO:710 on a ring-monotone cloud
    else if (key < Q.k3) { Q.k3 = key; assoc_update_cut<CORNER>(Q); }                         // LO:688 / LO:716
}
// ---- polar index primitives shared by the two searches ------------------------------------------------------------
struct RingBands { float elo[LL_MAX_RINGS], ehi[LL_MAX_RINGS]; int ordered; };
struct PolarQuery {
    float rho, qn, eq, frac, inv_w;  // horizontal range, range, elevation, position inside the own azimuth bin [0,1], bins per radian
    int b0, r0;                      // own azimuth bin; first ring whose band ends at or above the query's elevation
};
struct PolarReach { int ra, rb, kl, kr; };  // rings [ra, rb]; azimuth bins b0-kl .. b0+kr
#define NN_NONE ((u64)0x41C80000u << 32)    // (bits of 25.0f, index 0): only d2 < 25 beats it (LO:497 / LO:659)
__device__ __forceinline__ int azimuth_bin_frac(float x, float y, int NB, float& frac)
{
    const float phi = atan2f(y, x);  // same arithmetic as azimuth_bin()
    const float fb = (phi + 3.14159265f) * ((float)NB * 0.15915494f);
    int b = (int)floorf(fb);
    b = b < 0 ? 0 : (b >= NB ? NB - 1 : b);
    frac = fminf(fmaxf(fb - (float)b, 0.f), 1.f);
    return b;
}
// Largest angle (+ slack) between the query's direction and that of a point closer than bd, for a query `range` away
// from the apex: asin(bd / range), over-estimated by 1.0472 x (x <= 0.5) or (pi/2) x; 4 (> pi) = anywhere.
__device__ __forceinline__ float reach_angle(float bd, float range)
{
    const float x = __fdividef(bd, range);
    if (!(x < 1.f)) return 4.f;
    return x * (x < 0.5f ? 1.0472f : 1.5708f) + 2e-4f;
}
__device__ __forceinline__ void polar_query_init(PolarQuery& pq, float qx, float qy, float qz, int NB)
{
    pq.rho = sqrtf(qx * qx + qy * qy);
    pq.qn = sqrtf(pq.rho * pq.rho + qz * qz);
    pq.eq = atan2f(qz, pq.rho);   // elevation angle (the bands hold atanf of the tangents, widened by 1e-5)
    pq.b0 = azimuth_bin_frac(qx, qy, NB, pq.frac);
    pq.inv_w = (float)NB * 0.15915494f;
}
// azimuth part of the reach: bins at left offset k begin (frac + k - 1) bins away, at right offset k (k - frac) bins
__device__ __forceinline__ void polar_reach_bins(const PolarQuery& pq, int NB, float bd, int& kl, int& kr)
{
    const float ab = reach_angle(bd, pq.rho) * pq.inv_w;
    kl = ab >= (float)NB ? NB : (int)floorf(ab - pq.frac + 1.f);
    kr = ab >= (float)NB ? NB : (int)floorf(ab + pq.frac);
    kl = max(0, min(kl, NB / 2 - 1));   // all the way round: every bin once
    kr = max(0, min(kr, NB / 2));
}
__device__ __forceinline__ float best_dist(u64 best) { return sqrtf(__uint_as_float((unsigned)(best >> 32))) + 2e-3f; }
// ---- ring window ---------------------------------------------------------------------------------------------
// Ring-monotone *Last cloud: the serial loops of LO:504-553 / LO:668-721 visit exactly the points whose ring lies
// in [cring-2, cring+2].  Those rings are searched through the polar index: a point whose azimuth differs from the
// query's by D lies at least rho * sin(D) away (rho = horizontal range of the query), so bins are visited outward from
// the query's position inside its own bin until that bound exceeds the distances held (or 5 m, LO:29).  The index is
// bin-major (bucket = bin * R + ring): the five rings of one bin are ONE contiguous span of the sorted array.
// distance within which the window still has to look: the farther of the 2nd / 3rd point held, 5 m while one is open
template <bool CORNER>
__device__ __forceinline__ float assoc_window_dist(const AssocQuery& Q)
{
    const unsigned h2 = (unsigned)(Q.k2 >> 32), h3 = CORNER ? 0u : (unsigned)(Q.k3 >> 32);
    const unsigned h = min(max(h2, h3), D2_BITS_25);
    return sqrtf(__uint_as_float(h)) + 2e-3f;
}
// Per-thread part: up to dmax bins per side.  Returns false when the window needs more than that (the warp pass
// restarts it).
template <bool CORNER, typename ST>
__device__ __forceinline__ bool assoc_ring_window(AssocQuery& Q, const PolarQuery& pq, const ST& st, int NB, int R, int dmax, int* stat)
{
    const int r_lo = max(Q.cring - 2, 0), r_n = min(Q.cring + 2, R - 1) + 1 - r_lo;
    auto header = [&](int off, int& s, int& e) { s = st.header(pq.b0 + off, r_lo); e = st.header(pq.b0 + off, r_lo + r_n); };
    auto consider = [&](const float4 t) { assoc_consider<CORNER>(Q, t); };
    // own bin first: at long range a bin is wider than the distances in play and mostly settles the window alone; the
    // neighbours follow only as far as rho * sin(azimuth gap) stays inside the distances held
    int s0, e0, s1, e1;
    header(0, s0, e0);
    ASTAT(2, e0 - s0);
    for_spans3(st, s0, e0, 0, 0, 0, 0, consider);
    int kl, kr;
#pragma unroll 1
    for (int k = 1;; ++k) {
        polar_reach_bins(pq, NB, assoc_window_dist<CORNER>(Q), kl, kr);
        if (k > kl && k > kr) return true;
        if (k > dmax) return false;
        s0 = e0 = s1 = e1 = 0;
        if (k <= kl) header(-k, s0, e0);
        if (k <= kr) header(k, s1, e1);
        ASTAT(3, (e0 - s0) + (e1 - s1));
        for_spans3(st, s0, e0, s1, e1, 0, 0, consider);
    }
}
// Warp part: one query's whole ring window, the bins in reach shared out over the lanes (at most 16 per side and
// round, the reach being re-evaluated between rounds).
// `step` = bins per side of the first round (doubling up to 16): 16 for the queue's leftovers, which are wide by selection;
// 2 when the warp takes a fresh query (k_odom_assoc_direct), whose window mostly closes inside its own bin and the next.
template <bool CORNER>
__device__ __forceinline__ void assoc_ring_window_warp(AssocQuery& Q, const int* __restrict__ start, const float4* __restrict__ sorted, int NB, int R, int step = 16)
{
    const int lane = lane_id();
    PolarQuery pq;
    polar_query_init(pq, Q.qx, Q.qy, Q.qz, NB);
    const int r_lo = max(Q.cring - 2, 0), r_n = min(Q.cring + 2, R - 1) + 1 - r_lo;
    int doneL = 0, doneR = -1;   // left offsets 1..doneL and right offsets 0..doneR have been visited
    for (;; step = min(step * 2, 16)) {
        int kl, kr;
        polar_reach_bins(pq, NB, assoc_window_dist<CORNER>(Q), kl, kr);
        if (doneL >= kl && doneR >= kr) break;
        const int idx = lane >> 1;
        int beg = 0, cnt = 0, off = 0;
        bool on = false;
        if (lane & 1) { const int k = doneL + 1 + idx; on = idx < step && k <= kl; off = -k; }
        else { const int k = doneR + 1 + idx; on = idx < step && k <= kr; off = k; }
        if (on) {
            const int* h = start + ((pq.b0 + off) & (NB - 1)) * R + r_lo;
            beg = __ldg(h);
            cnt = __ldg(h + r_n) - beg;
        }
        grid_stream_ranges(sorted, beg, cnt, [&](const float4 t) { assoc_consider<CORNER>(Q, t); });
        doneL += step; doneR += step;
        Q.k2 = warp_min_u64(Q.k2);
        if (!CORNER) Q.k3 = warp_min_u64(Q.k3);
        assoc_update_cut<CORNER>(Q);
    }
}
// The literal scan loops (LO:504-553 / LO:668-721) for clouds that are not ring-sorted (legal on the topic, never
// produced by scanRegistration); the warp evaluates 32 candidates per ballot, lane order = visit order.
template <bool CORNER>
__device__ __forceinline__ void assoc_literal_walk_warp(AssocQuery& Q, const float4* __restrict__ last)
{
    const int lane = lane_id();
    const int closest = Q.closest, cring = Q.cring, n = Q.n;
    bool stop = false;
    for (int j0 = closest + 1; j0 < n && !stop; j0 += 32) {  // increasing scan line
        const int j = j0 + lane;
        bool brk = false;
        u64 c2 = ~0ull, c3 = ~0ull;
        if (j < n) {
            const float4 t = last[j];
            const int rj = (int)t.w;
            brk = rj > cring + 2;   // int ring vs cring + 2.5 (LO:511)
            const float d2 = sqdist3(t.x, t.y, t.z, Q.qx, Q.qy, Q.qz);
            const u64 key = ((u64)__float_as_uint(d2) << 32) | (unsigned)(j - (closest + 1));
            if (!brk && d2 < 25.0f) {
                if (CORNER) { if (!(rj <= cring)) c2 = key; }  // LO:507: same scan line -> continue
                else if (rj <= cring) c2 = key;                // LO:682
                else c3 = key;                                 // LO:688
            }
        }
        const unsigned bm = __ballot_sync(LL_FULL_MASK, brk);
        const int fb = bm ? __ffs(bm) - 1 : 32;
        if (lane < fb) { if (c2 < Q.k2) Q.k2 = c2; if (c3 < Q.k3) Q.k3 = c3; }
        if (bm) stop = true;
    }
    stop = false;
    for (int j0 = closest - 1; j0 >= 0 && !stop; j0 -= 32) {  // decreasing scan line; visit order continues after the up-scan
        const int j = j0 - lane;
        bool brk = false;
        u64 c2 = ~0ull, c3 = ~0ull;
        if (j >= 0) {
            const float4 t = last[j];
            const int rj = (int)t.w;
            brk = rj < cring - 2;   // LO:537
            const float d2 = sqdist3(t.x, t.y, t.z, Q.qx, Q.qy, Q.qz);
            const u64 key = ((u64)__float_as_uint(d2) << 32) | ((unsigned)n + (unsigned)(closest - 1 - j));
            if (!brk && d2 < 25.0f) {
                if (CORNER) { if (!(rj >= cring)) c2 = key; }
                else if (rj >= cring) c2 = key;
                else c3 = key;
            }
        }
        const unsigned bm = __ballot_sync(LL_FULL_MASK, brk);
        const int fb = bm ? __ffs(bm) - 1 : 32;
        if (lane < fb) { if (c2 < Q.k2) Q.k2 = c2; if (c3 < Q.k3) Q.k3 = c3; }
        if (bm) stop = true;
    }
    Q.k2 = warp_min_u64(Q.k2);
    Q.k3 = warp_min_u64(Q.k3);
}
// ---- exact 1-NN over the polar index ---------------------------------------------------------------------------
// per-thread version, bands in shared memory
__device__ __forceinline__ PolarReach polar_reach(const RingBands& S, const PolarQuery& pq, int NB, int R, u64 best)
{
    const float bd = best_dist(best);
    PolarReach W;
    const float g = reach_angle(bd, pq.qn);
    int ra = pq.r0, rb = pq.r0;
    if (!S.ordered) { ra = 0; rb = R - 1; }
    else {
        while (ra > 0 && pq.eq - S.ehi[ra - 1] <= g) --ra;
        while (rb < R - 1 && S.elo[rb + 1] - pq.eq <= g) ++rb;
    }
    W.ra = ra; W.rb = rb;
    polar_reach_bins(pq, NB, bd, W.kl, W.kr);
    return W;
}
__device__ __forceinline__ void nn_consider(u64& best, float qx, float qy, float qz, const float4 t)
{
    const float d2 = sqdist3(qx, qy, qz, t.x, t.y, t.z);
    const u64 key = ((u64)__float_as_uint(d2) << 32) | ((unsigned)__float_as_int(t.w) & 0xFFFFFFu);
    if (key < best) best = key;   // NaN distances have bits above 25.0f's and never win
}
// the same, remembering the winner's ring (bits 24..31 of .w)
__device__ __forceinline__ void nn_consider_ring(u64& best, int& best_ring, float qx, float qy, float qz, const float4 t)
{
    const float d2 = sqdist3(qx, qy, qz, t.x, t.y, t.z);
    if (__float_as_uint(d2) > (unsigned)(best >> 32)) return;   // most candidates: one 32-bit compare (equal bits go on to the index tie-break)
    const unsigned bits = (unsigned)__float_as_int(t.w);
    const u64 key = ((u64)__float_as_uint(d2) << 32) | (bits & 0xFFFFFFu);
    if (key < best) { best = key; best_ring = (int)(bits >> 24); }
}
// Per-thread search.  Returns true when `best` is final; false = too wide for one thread (best = what was found so
// far, the warp pass restarts from it).
template <typename ST>
__device__ __forceinline__ bool polar_nearest_thread(const RingBands& S, const ST& st, int NB, int R, float qx, float qy, float qz, int kmax, PolarQuery& pq, u64& best,
                                                     int& best_ring, int& reach_buckets, int* stat)
{
    {
        int lo = 0, hi = R;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (S.ehi[mid] < pq.eq) lo = mid + 1; else hi = mid; }
        pq.r0 = min(lo, R - 1);
    }
    auto consider = [&](const float4 t) { nn_consider_ring(best, best_ring, qx, qy, qz, t); };
    auto span = [&](int off, int ra, int rb, int& s, int& e) { s = st.header(pq.b0 + off, ra); e = st.header(pq.b0 + off, rb + 1); };
    // seed: rings r0-1 .. r0+1 of the own bin (one contiguous span).  At long range - where the buckets are fullest - a
    // bin is wider than the match distance, so the seed usually settles the search; the rest of the reach follows below.
    const int sa = max(pq.r0 - 1, 0), sb = min(pq.r0 + 1, R - 1);
    int s0, e0, s1, e1;
    span(0, sa, sb, s0, e0);
    ASTAT(0, e0 - s0);
    for_spans3(st, s0, e0, 0, 0, 0, 0, consider);
    PolarReach W = polar_reach(S, pq, NB, R, best);
    if (W.ra >= sa && W.rb <= sb && W.kl == 0 && W.kr == 0) return true;   // nothing closer can lie outside the seed
    if (W.rb - W.ra > 16 || max(W.kl, W.kr) > kmax) {
        // the seed found nothing close: its two neighbour bins first (one stream) - they usually pull the reach back
        // inside what one thread walks; only what is still too wide afterwards goes to the warp pass
        int s2, e2;
        span(-1, sa, sb, s0, e0); span(1, sa, sb, s2, e2);
        ASTAT(1, (e0 - s0) + (e2 - s2));
        for_spans3(st, s0, e0, s2, e2, 0, 0, consider);
        W = polar_reach(S, pq, NB, R, best);
        if (W.ra >= sa && W.rb <= sb && W.kl <= 1 && W.kr <= 1) return true;
        reach_buckets = (W.kl + W.kr + 1) * (W.rb - W.ra + 1);
        if (W.rb - W.ra > 16 || max(W.kl, W.kr) > kmax) return false;
    }
    // the rings of the own bin the seed left out, then outward over the bins in reach; the reach only shrinks as the best improves
    u64 seen = best;
    if (W.ra < sa || W.rb > sb) {
        s0 = e0 = s1 = e1 = 0;
        if (W.ra < sa) span(0, W.ra, sa - 1, s0, e0);
        if (W.rb > sb) span(0, sb + 1, W.rb, s1, e1);
        ASTAT(1, (e0 - s0) + (e1 - s1));
        for_spans3(st, s0, e0, s1, e1, 0, 0, consider);
    }
#pragma unroll 1
    for (int k = 1;; ++k) {
        if (best != seen) { W = polar_reach(S, pq, NB, R, best); seen = best; }
        if (k > W.kl && k > W.kr) return true;
        s0 = e0 = s1 = e1 = 0;
        if (k <= W.kl) span(-k, W.ra, W.rb, s0, e0);
        if (k <= W.kr) span(k, W.ra, W.rb, s1, e1);
        ASTAT(1, (e0 - s0) + (e1 - s1));
        for_spans3(st, s0, e0, s1, e1, 0, 0, consider);
    }
}
// Warp version (one query per warp; bands read from global memory): every step streams the rings in reach of up to
// 16 bins per side.  `best` (warp-uniform) may carry what the thread pass found.
__device__ __forceinline__ u64 polar_nearest_warp(const float* __restrict__ bands, const int* __restrict__ start, const float4* __restrict__ sorted, int NB, int R,
                                                  float qx, float qy, float qz, u64 best, int wk = 8, int wr = 16)
{
    const int lane = lane_id();
    PolarQuery pq;
    polar_query_init(pq, qx, qy, qz, NB);
    // this lane's two rings
    const float elo0 = lane < R ? __ldg(bands + lane) : INFINITY, ehi0 = lane < R ? __ldg(bands + LL_MAX_RINGS + lane) : INFINITY;
    const float elo1 = lane + 32 < R ? __ldg(bands + lane + 32) : INFINITY, ehi1 = lane + 32 < R ? __ldg(bands + LL_MAX_RINGS + lane + 32) : INFINITY;
    const int ordered = __ldg(reinterpret_cast<const int*>(bands) + 2 * LL_MAX_RINGS);
    pq.r0 = min(__popc(__ballot_sync(LL_FULL_MASK, ehi0 < pq.eq)) + __popc(__ballot_sync(LL_FULL_MASK, ehi1 < pq.eq)), R - 1);
    u64 lb = best;
    // The window (bins b0-wk .. b0+wk, rings r0-wr .. r0+wr) doubles every round and is clipped to the reach of the best
    // found so far; the search ends when a round's window covered the whole reach it was clipped to.  The cost is
    // that of the final window (the earlier ones add a third), not that of the 5 m reach a query starts with.
    int pra = 1 << 20, prb = -1, pkl = -1, pkr = -1;   // rings / bins the previous round streamed (nothing yet)
    // first window: 8 x 16 by default - what the thread pass hands over nearly always resolves inside it; a fresh query
    // (k_odom_assoc_direct) starts with its own bin and ring and their neighbours
    for (;;) {
        const float bd = best_dist(best);
        const float g = reach_angle(bd, pq.qn);
        int ra = 0, rb = R - 1, kl, kr;
        if (ordered) {
            // ring r is in reach when its band comes within g of the query's elevation; ordered bands: a contiguous range
            const bool in0 = lane < R && (lane >= pq.r0 ? elo0 - pq.eq <= g : pq.eq - ehi0 <= g);
            const bool in1 = lane + 32 < R && (lane + 32 >= pq.r0 ? elo1 - pq.eq <= g : pq.eq - ehi1 <= g);
            const u64 m = ((u64)__ballot_sync(LL_FULL_MASK, in1) << 32) | __ballot_sync(LL_FULL_MASK, in0);
            // the run of set bits around r0
            const u64 below = ~m & ((1ull << pq.r0) - 1ull);              // rings < r0 out of reach
            ra = below ? 64 - __clzll(below) : 0;
            const u64 above = pq.r0 >= 63 ? 0ull : (~m >> (pq.r0 + 1));   // rings > r0 out of reach, bit 0 = r0 + 1
            rb = above ? pq.r0 + __ffsll(above) - 1 : 63;
            rb = min(rb, R - 1);
        }
        polar_reach_bins(pq, NB, bd, kl, kr);
        if (ra >= pra && rb <= prb && kl <= pkl && kr <= pkr) break;   // the last window already covered this reach
        const int cra = max(ra, pq.r0 - wr), crb = min(rb, pq.r0 + wr), ckl = min(kl, wk), ckr = min(kr, wk);
        pra = cra; prb = crb; pkl = ckl; pkr = ckr;
        for (int o0 = -ckl; o0 <= ckr; o0 += 32) {
            const int off = o0 + lane;
            int beg = 0, cnt = 0;
            if (off <= ckr) {
                const int* h = start + ((pq.b0 + off) & (NB - 1)) * R;
                beg = __ldg(h + cra);
                cnt = __ldg(h + crb + 1) - beg;
            }
            grid_stream_ranges(sorted, beg, cnt, [&](const float4 t) { nn_consider(lb, qx, qy, qz, t); });
        }
        best = warp_min_u64(lb);
        if (cra == ra && crb == rb && ckl == kl && ckr == kr) break;   // the reach only shrinks from here
        wk <<= 1; wr <<= 1;
    }
    return best;
}
// view of one lane's polar index
__device__ __forceinline__ GridView assoc_view(const KnnGrid& G, int b)
{
    GridView gv;
    gv.start = G.start + (size_t)b * (G.T + 1);
    gv.sorted = G.sorted + (size_t)b * G.cap;
    gv.Tmask = G.T - 1;
    gv.h = G.h;
    gv.inv_h = G.inv_h;
    return gv;
}
template <bool CORNER>
__device__ __forceinline__ void assoc_store(const OdomParams& P, int b, int i, const AssocQuery& Q)
{
    auto decode = [&](u64 k) -> int {
        if (k == ~0ull) return -1;
        const unsigned rk = (unsigned)k;
        return rk >= (unsigned)Q.n ? Q.closest - 1 - (int)(rk - (unsigned)Q.n) : Q.closest + 1 + (int)rk;
    };
    const int ind2 = Q.closest >= 0 ? decode(Q.k2) : -1, ind3 = Q.closest >= 0 ? decode(Q.k3) : -1;
    if (CORNER) {
        int* o = P.corner_assoc + ((size_t)b * P.R * LL_SHARP_PER_RING + i) * 2;
        *reinterpret_cast<int2*>(o) = make_int2(ind2 >= 0 ? Q.closest : -1, ind2);  // LO:556
    } else {
        int* o = P.plane_assoc + ((size_t)b * P.R * LL_FLAT_PER_RING + i) * 4;
        const bool ok = ind2 >= 0 && ind3 >= 0;  // LO:723
        *reinterpret_cast<int4*>(o) = ok ? make_int4(Q.closest, ind2, ind3, 0) : make_int4(-1, -1, -1, 0);
    }
}
// LO:81-84 / LO:569-573 / LO:739-743 with DISTORTION 1: (intensity - int(intensity)) / SCAN_PERIOD - a float difference divided in double
__device__ __forceinline__ double point_ratio(const float4 p) { return (double)(p.w - (float)(int)p.w) / 0.1; }
template <bool CORNER>
__device__ __forceinline__ void assoc_query_init(const OdomParams& P, const LaneState& L, int b, int i, AssocQuery& Q)
{
    const float4 p = CORNER ? P.sharp[(size_t)b * P.R * LL_SHARP_PER_RING + i] : P.flat[(size_t)b * P.R * LL_FLAT_PER_RING + i];
    double sx, sy, sz;
    if (P.distortion) {
        // TransformToStart, LO:77-95 with DISTORTION 1: s = fraction of the intensity / SCAN_PERIOD, q_s = Identity.slerp(s, q), t_s = s t
        const double s = point_ratio(p);
        double qs[4];
        lm_identity_slerp<double>(s, L.para_q, qs);
        quat_rotate(qs, (double)p.x, (double)p.y, (double)p.z, sx, sy, sz);
        Q.qx = (float)(sx + s * L.para_t[0]); Q.qy = (float)(sy + s * L.para_t[1]); Q.qz = (float)(sz + s * L.para_t[2]);
    } else {
        // TransformToStart, LO:77-95 with DISTORTION 0: slerp(1, q) = +-q, same rotation bit for bit
        quat_rotate(L.para_q, (double)p.x, (double)p.y, (double)p.z, sx, sy, sz);
        Q.qx = (float)(sx + L.para_t[0]); Q.qy = (float)(sy + L.para_t[1]); Q.qz = (float)(sz + L.para_t[2]);
    }
    Q.k2 = ~0ull; Q.k3 = ~0ull; Q.closest = -1; Q.cring = 0; Q.cut = D2_BITS_25 - 1u;
    Q.n = CORNER ? L.n_last_corner : L.n_last_surf;
}
// queue entry of a query the per-thread pass did not finish: x = lane, y = feature index | plane << 30 | open << 29;
// open (nearest neighbour not final): z, w = best key so far (d2 bits, index; NN_NONE = nothing within 5 m yet)
// otherwise: z = closest point, w = -1 (ring window, restarted by the warp) or -2 (literal walk)
#define ASSOC_PLANE_BIT (1 << 30)
#define ASSOC_OPEN_BIT (1 << 29)

// Per-thread pass.  Writes the correspondences of the queries it resolves, queues the others for k_odom_assoc_heavy.
// Q / pq arrive initialised (query point, own azimuth bin); st = where the polar index is read from.
template <bool CORNER, typename ST>
__device__ __forceinline__ void assoc_thread_pass(const OdomParams& P, const RingBands& S, const ST& st, const LaneState& L, int b, int i, bool active, int dmax, int kmax,
                                                  AssocQuery& Q, PolarQuery& pq)
{
    const int lane = lane_id();
    int4 entry = make_int4(b, CORNER ? i : (i | ASSOC_PLANE_BIT), -1, -1);
    bool heavy = false, is_long = false;
    if (active) {
        const int NB = CORNER ? P.az_bins_corner : P.az_bins_surf;
        u64 best = NN_NONE;
        int best_ring = 0;
#ifdef LL_ASSOC_STATS
        int stat[4] = {0, 0, 0, 0};
#else
        int* stat = nullptr;
#endif
        if (Q.n > 0) {
            int reach_buckets = 0;
            heavy = !polar_nearest_thread(S, st, NB, P.R, Q.qx, Q.qy, Q.qz, kmax, pq, best, best_ring, reach_buckets, stat);
            is_long = heavy && reach_buckets > 200;   // these take tens of microseconds each: they must start first
            if (heavy) { entry.y |= ASSOC_OPEN_BIT; entry.z = (int)(unsigned)(best >> 32); entry.w = (int)(unsigned)best; }
        }
        if (!heavy && best != NN_NONE) {  // d2 < 25: LO:497 / LO:659
            Q.closest = (int)(unsigned)best;
            Q.cring = best_ring;               // int(intensity) of the closest point (LO:500 / LO:664): the index carries it for ring-monotone clouds
            int pending = -2;                  // clouds that are not ring-sorted: literal walk
            if (CORNER ? L.mono_corner : L.mono_surf) {
                pending = assoc_ring_window<CORNER>(Q, pq, st, NB, P.R, dmax, stat) ? -3 : -1;
                heavy = pending == -1;
            } else {
                heavy = true;
            }
            entry.z = Q.closest;
            entry.w = pending;
        }
        if (!heavy) assoc_store<CORNER>(P, b, i, Q);
#ifdef LL_ASSOC_STATS
        {
            const int cls = (CORNER ? 0 : 16) + (pq.rho < 10.f ? 0 : (pq.rho < 20.f ? 5 : 10));
            atomicAdd(&g_assoc_stats[cls], 1ull);
            for (int k = 0; k < 4; ++k) atomicAdd(&g_assoc_stats[cls + 1 + k], (unsigned long long)stat[k]);
            if (heavy) atomicAdd(&g_assoc_stats[CORNER ? 15 : 31], 1ull);
        }
#endif
    }
    // the long entries are stored from the front of the queue, the others from its back; the warp pass pops front to back
    // (longest-first: the kernel cannot end before its longest entry does, so that one must not start last)
#pragma unroll
    for (int cls = 0; cls < 2; ++cls) {
        const bool mine = heavy && (is_long == (cls == 0));
        const unsigned hm = __ballot_sync(LL_FULL_MASK, mine);
        if (hm) {
            const int leader = __ffs(hm) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(P.queue_n + (cls == 0 ? 0 : 8) + P.outer, __popc(hm));
            base = __shfl_sync(LL_FULL_MASK, base, leader);
            const int pos = base + __popc(hm & ((1u << lane) - 1u));
            if (mine) P.queue[cls == 0 ? pos : P.queue_cap - 1 - pos] = entry;   // long + short <= number of queries <= queue_cap
        }
    }
}
template <bool CORNER>
__device__ __forceinline__ void assoc_thread_query_global(const OdomParams& P, const RingBands& S, const LaneState& L, int b, int i, bool active, int dmax, int kmax)
{
    AssocQuery Q;
    Q.qx = Q.qy = Q.qz = 0.f; Q.k2 = Q.k3 = ~0ull; Q.closest = -1; Q.cring = 0; Q.n = 0; Q.cut = D2_BITS_25 - 1u;
    PolarQuery pq;
    pq.rho = pq.qn = pq.eq = pq.frac = pq.inv_w = 0.f; pq.b0 = pq.r0 = 0;
    const KnnGrid& G = CORNER ? P.ac : P.as_;
    const int NB = CORNER ? P.az_bins_corner : P.az_bins_surf;
    StageGlobal st;
    st.start = G.start + (size_t)b * (G.T + 1); st.sorted = G.sorted + (size_t)b * G.cap; st.NB = NB; st.R = P.R;
    if (active) {
        assoc_query_init<CORNER>(P, L, b, i, Q);
        polar_query_init(pq, Q.qx, Q.qy, Q.qz, NB);
    }
    assoc_thread_pass<CORNER>(P, S, st, L, b, i, active, dmax, kmax, Q, pq);
}
// grid: (ceil(R*12 / T) + ceil(R*24 / T), B): corner queries and plane queries never share a block
template <int MINB>
__global__ void __launch_bounds__(ASSOC_THREADS, MINB) k_odom_assoc(OdomParams P, int corner_blocks, int dmax, int kmax)
{
    __shared__ RingBands S;
    const int b = blockIdx.y;
    const LaneState& L = P.lane[b];
    if (!L.inited || L.err) return;
    const bool corner = (int)blockIdx.x < corner_blocks;
    const int i = (corner ? blockIdx.x : blockIdx.x - corner_blocks) * (int)blockDim.x + threadIdx.x;   // blockDim.x <= ASSOC_THREADS
    const int n = corner ? L.n_sharp : L.n_flat;
    if ((int)(i - threadIdx.x) >= n) return;   // whole block idle
    {
        const float* bd = P.bands[corner ? 0 : 1] + (size_t)b * BAND_STRIDE;
        if (threadIdx.x < P.R) { S.elo[threadIdx.x] = __ldg(bd + threadIdx.x); S.ehi[threadIdx.x] = __ldg(bd + LL_MAX_RINGS + threadIdx.x); }
        if (threadIdx.x == 0) S.ordered = __ldg(reinterpret_cast<const int*>(bd) + 2 * LL_MAX_RINGS);
        __syncthreads();
    }
    if ((i & ~31) >= n) return;   // whole warp idle
    if (corner) assoc_thread_query_global<true>(P, S, L, b, i, i < n, dmax, kmax);
    else assoc_thread_query_global<false>(P, S, L, b, i, i < n, dmax, kmax);
}

// ------------------------------------------------------------------------------------------------------
// Slab form of the thread pass (the default): the polar index is bin-major, so the buckets of a range of azimuth bins
// - all rings - are ONE contiguous span of the bucket-sorted array.  k_odom_queries transforms the lane's feature
// points once per outer iteration (TransformToStart) and groups them by azimuth slab; one CTA of k_odom_assoc_slab
// then takes one (slab, lane): it pulls the slab's span plus `halo` bins on each side into shared memory with TMA bulk
// copies (cp.async.bulk + mbarrier; two pieces when the range wraps past bin NB - 1), rewrites the bucket headers of
// those bins to positions inside the staged array, and runs the same per-thread searches as above out of shared
// memory: every target point crosses the memory system once per slab as part of a bulk copy instead of once per
// visiting query as a dependent 16-byte load.  A thread walks at most `halo` bins to either side (kmax = dmax = halo),
// so it never leaves the stage; what needs more goes to the warp pass as before.  A span that does not fit the
// CTA's shared memory (a direction crowded with points) is searched in global memory by the same code.
// ------------------------------------------------------------------------------------------------------
struct SlabParams {
    float4* qa;        // [B][R*36] x, y, z of the transformed query, w = feature index bits
    float4* qb;        // [B][R*36] elevation angle, position inside the azimuth bin, own bin (int bits), unused
    int* qstart;       // [B][ns_corner + ns_surf + 2] slab offsets into qa / qb (corner slabs, then surf slabs from R*12)
    int ns_corner, ns_surf;       // slabs per table
    int slab_corner, slab_surf;   // bins per slab
    int halo_corner, halo_surf;
    int pcap;                     // points the stage of one CTA holds
};
#define QPREP_THREADS 1024
__global__ void __launch_bounds__(QPREP_THREADS) k_odom_queries(OdomParams P, SlabParams Sp)
{
    __shared__ int cnt[2][130];
    const int b = blockIdx.x, tid = threadIdx.x;
    const LaneState& L = P.lane[b];
    const int nsl[2] = {Sp.ns_corner, Sp.ns_surf};
    int* qs = Sp.qstart + (size_t)b * (Sp.ns_corner + Sp.ns_surf + 2);
    for (int k = tid; k < 2 * 130; k += QPREP_THREADS) (&cnt[0][0])[k] = 0;
    __syncthreads();
    const bool live = L.inited && !L.err;
    const int ns = live ? L.n_sharp : 0, nf = live ? L.n_flat : 0;
    const int maxc = P.R * LL_SHARP_PER_RING;
    // work items: corner i = tid (R*12 <= 768), planes i = tid, tid + 1024 (R*24 <= 1536)
    float4 a[3], c[3];
    int slab[3], slot[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
        const bool corner = u == 0;
        const int i = corner ? tid : tid + (u - 1) * QPREP_THREADS;
        slab[u] = -1;
        if (i < (corner ? ns : nf)) {
            AssocQuery Q;
            if (corner) assoc_query_init<true>(P, L, b, i, Q); else assoc_query_init<false>(P, L, b, i, Q);
            PolarQuery pq;
            const int NB = corner ? P.az_bins_corner : P.az_bins_surf;
            polar_query_init(pq, Q.qx, Q.qy, Q.qz, NB);
            a[u] = make_float4(Q.qx, Q.qy, Q.qz, __int_as_float(i));
            c[u] = make_float4(pq.eq, pq.frac, __int_as_float(pq.b0), 0.f);
            slab[u] = pq.b0 / (corner ? Sp.slab_corner : Sp.slab_surf);
            slot[u] = atomicAdd(&cnt[corner ? 0 : 1][slab[u]], 1);
        }
    }
    __syncthreads();
    if (tid < 64) {   // exclusive scans of the two count arrays (<= 128 slabs each), one warp each
        const int t = tid >> 5, ln = tid & 31;
        int run = 0;
        for (int k0 = 0; k0 < nsl[t]; k0 += 32) {
            const int k = k0 + ln;
            const int v = k < nsl[t] ? cnt[t][k] : 0;
            int incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(LL_FULL_MASK, incl, d); if (ln >= d) incl += o; }
            if (k < nsl[t]) cnt[t][k] = run + incl - v;
            run += __shfl_sync(LL_FULL_MASK, incl, 31);
        }
        if (ln == 0) cnt[t][nsl[t]] = run;
    }
    __syncthreads();
    for (int k = tid; k <= Sp.ns_corner; k += QPREP_THREADS) qs[k] = cnt[0][k];
    for (int k = tid; k <= Sp.ns_surf; k += QPREP_THREADS) qs[Sp.ns_corner + 1 + k] = maxc + cnt[1][k];
    float4* qa = Sp.qa + (size_t)b * P.R * (LL_SHARP_PER_RING + LL_FLAT_PER_RING);
    float4* qb = Sp.qb + (size_t)b * P.R * (LL_SHARP_PER_RING + LL_FLAT_PER_RING);
#pragma unroll
    for (int u = 0; u < 3; ++u) {
        if (slab[u] < 0) continue;
        const int pos = (u == 0 ? 0 : maxc) + cnt[u == 0 ? 0 : 1][slab[u]] + slot[u];
        qa[pos] = a[u];
        qb[pos] = c[u];
    }
}

template <bool CORNER, typename ST>
__device__ __forceinline__ void slab_queries(const OdomParams& P, const SlabParams& Sp, const RingBands& S, const ST& st, const LaneState& L, int b, int q0, int q1, int halo)
{
    const int NB = CORNER ? P.az_bins_corner : P.az_bins_surf;
    const float4* qa = Sp.qa + (size_t)b * P.R * (LL_SHARP_PER_RING + LL_FLAT_PER_RING);
    const float4* qb = Sp.qb + (size_t)b * P.R * (LL_SHARP_PER_RING + LL_FLAT_PER_RING);
    const int n_last = CORNER ? L.n_last_corner : L.n_last_surf;
    for (int base = q0 + (int)(threadIdx.x & ~31u); base < q1; base += (int)blockDim.x) {   // warps stay whole: the queue push uses ballots
        const int q = base + lane_id();
        const bool active = q < q1;
        AssocQuery Q;
        Q.qx = Q.qy = Q.qz = 0.f; Q.k2 = Q.k3 = ~0ull; Q.closest = -1; Q.cring = 0; Q.n = 0; Q.cut = D2_BITS_25 - 1u;
        PolarQuery pq;
        pq.rho = pq.qn = pq.eq = pq.frac = pq.inv_w = 0.f; pq.b0 = pq.r0 = 0;
        int i = 0;
        if (active) {
            const float4 a = __ldg(qa + q), c = __ldg(qb + q);
            Q.qx = a.x; Q.qy = a.y; Q.qz = a.z; Q.n = n_last;
            i = __float_as_int(a.w);
            pq.rho = sqrtf(a.x * a.x + a.y * a.y);
            pq.qn = sqrtf(pq.rho * pq.rho + a.z * a.z);
            pq.eq = c.x; pq.frac = c.y; pq.b0 = __float_as_int(c.z);
            pq.inv_w = (float)NB * 0.15915494f;
        }
        assoc_thread_pass<CORNER>(P, S, st, L, b, i, active, halo, halo, Q, pq);
    }
}

#define SLAB_THREADS 128
template <bool CORNER>
__device__ __forceinline__ void slab_cta(const OdomParams& P, const SlabParams& Sp, RingBands& S, unsigned char* smem, uint64_t* bar, const LaneState& L, int b, int slab)
{
    const int tid = threadIdx.x;
    const int NB = CORNER ? P.az_bins_corner : P.az_bins_surf, R = P.R;
    const int sl = CORNER ? Sp.slab_corner : Sp.slab_surf, halo = CORNER ? Sp.halo_corner : Sp.halo_surf;
    const int* qs = Sp.qstart + (size_t)b * (Sp.ns_corner + Sp.ns_surf + 2) + (CORNER ? 0 : Sp.ns_corner + 1);
    const int q0 = __ldg(qs + slab), q1 = __ldg(qs + slab + 1);
    if (q1 <= q0) return;
    const KnnGrid& G = CORNER ? P.ac : P.as_;
    const int* start = G.start + (size_t)b * (G.T + 1);
    const float4* sorted = G.sorted + (size_t)b * G.cap;
    // staged bins: [slab * sl - halo, slab * sl + sl + halo) modulo NB, or the whole table when that is all of it
    int nb = sl + 2 * halo, lo = (slab * sl - halo) & (NB - 1);
    if (nb >= NB) { nb = NB; lo = 0; }
    const int hiA = min(lo + nb, NB), nbB = lo + nb - hiA;          // piece A: bins [lo, hiA), piece B (wrap): bins [0, nbB)
    const int n_total = __ldg(start + NB * R);
    const int pA0 = __ldg(start + lo * R), pA1 = __ldg(start + hiA * R), pB1 = nbB > 0 ? __ldg(start + nbB * R) : 0;
    const int lenA = pA1 - pA0, lenB = pB1;
    int* hdr = reinterpret_cast<int*>(smem);
    float4* pts = reinterpret_cast<float4*>(smem + (((size_t)(sl + 2 * halo) * R + 1) * 4 + 15) / 16 * 16);
    {
        const float* bd = P.bands[CORNER ? 0 : 1] + (size_t)b * BAND_STRIDE;
        if (tid < R) { S.elo[tid] = __ldg(bd + tid); S.ehi[tid] = __ldg(bd + LL_MAX_RINGS + tid); }
        if (tid == 0) S.ordered = __ldg(reinterpret_cast<const int*>(bd) + 2 * LL_MAX_RINGS);
    }
    if (lenA + lenB > Sp.pcap) {   // does not fit: the same searches out of global memory
        __syncthreads();
        StageGlobal st;
        st.start = start; st.sorted = sorted; st.NB = NB; st.R = R;
        slab_queries<CORNER>(P, Sp, S, st, L, b, q0, q1, halo);
        return;
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        const uint32_t bytes = (uint32_t)(lenA + lenB) * 16u;
        if (bytes) {
            mbar_expect_tx(bar, bytes);
            if (lenA) tma_bulk_copy(pts, sorted + pA0, (uint32_t)lenA * 16u, bar);
            if (lenB) tma_bulk_copy(pts + lenA, sorted, (uint32_t)lenB * 16u, bar);
        }
    }
    // headers of the staged bins, rewritten to positions inside `pts` (piece B sits behind piece A)
    for (int h = tid; h <= nb * R; h += SLAB_THREADS) {
        const int g = lo * R + h;   // header index on the unwrapped bin axis
        hdr[h] = g <= NB * R ? __ldg(start + g) - pA0 : __ldg(start + (g - NB * R)) + (n_total - pA0);
    }
    __syncthreads();   // headers and bands visible, mbarrier initialised
    if (lenA + lenB > 0) mbar_wait(bar, 0);
    StageShared st;
    st.hdr = smem_u32(hdr); st.pts = smem_u32(pts); st.bin0 = lo; st.NB = NB; st.R = R;
    slab_queries<CORNER>(P, Sp, S, st, L, b, q0, q1, halo);
}
template <int MINB>
__global__ void __launch_bounds__(SLAB_THREADS, MINB) k_odom_assoc_slab(OdomParams P, SlabParams Sp)
{
    extern __shared__ __align__(128) unsigned char slab_smem[];
    __shared__ RingBands S;
    __shared__ __align__(8) uint64_t bar;
    const int b = blockIdx.y;
    const LaneState& L = P.lane[b];
    if (!L.inited || L.err) return;
    // surf slabs first: they carry the larger stages
    if ((int)blockIdx.x < Sp.ns_surf) slab_cta<false>(P, Sp, S, slab_smem, &bar, L, b, blockIdx.x);
    else slab_cta<true>(P, Sp, S, slab_smem, &bar, L, b, blockIdx.x - Sp.ns_surf);
}

// One WARP per queued query, the queue spread over a fixed grid: the long searches (no target nearby, wide ring
// windows, literal walks) run side by side instead of serialising inside the warp that met them.
template <bool CORNER>
__device__ __forceinline__ void assoc_warp_query(const OdomParams& P, int b, int i, bool open, int ez, int ew, bool fresh = false)
{
    const LaneState& L = P.lane[b];
    const int lane = lane_id();
    AssocQuery Q;
    assoc_query_init<CORNER>(P, L, b, i, Q);
    const float4* last = CORNER ? P.lsharp[last_slot(L)] + (size_t)b * P.R * LL_LSHARP_PER_RING : P.lflat[last_slot(L)] + (size_t)b * P.Nmax;
    const GridView av = assoc_view(CORNER ? P.ac : P.as_, b);
    const int NB = CORNER ? P.az_bins_corner : P.az_bins_surf;
    int closest = open ? -1 : ez, mode = open ? -1 : ew;
    if (open && Q.n > 0) {
        // 1-NN (kdtree*Last->nearestKSearch(pointSel, 1, ...), LO:494 / LO:656) continued from the thread pass's best
        const u64 best = polar_nearest_warp(P.bands[CORNER ? 0 : 1] + (size_t)b * BAND_STRIDE, av.start, av.sorted, NB, P.R, Q.qx, Q.qy, Q.qz,
                                            ((u64)(unsigned)ez << 32) | (unsigned)ew, fresh ? 1 : 8, fresh ? 1 : 16);
        if (best != NN_NONE) {  // d2 < 25: LO:497 / LO:659
            closest = (int)(unsigned)best;
            mode = (CORNER ? L.mono_corner : L.mono_surf) ? -1 : -2;
        }
    }
    if (closest >= 0) {
        Q.closest = closest;
        Q.cring = (int)last[closest].w;  // int(intensity), LO:500 / LO:664
        if (mode == -2) assoc_literal_walk_warp<CORNER>(Q, last);
        else assoc_ring_window_warp<CORNER>(Q, av.start, av.sorted, NB, P.R, fresh ? 2 : 16);
    }
    if (lane == 0) assoc_store<CORNER>(P, b, i, Q);
}
__global__ void __launch_bounds__(256, 4) k_odom_assoc_heavy(OdomParams P)
{
    const int n_long = P.queue_n[P.outer], n = n_long + P.queue_n[8 + P.outer];
    int* head = P.queue_n + 4 + P.outer;  // entries are popped one at a time: their costs differ by orders of magnitude
    for (;;) {
        int e = 0;
        if (lane_id() == 0) e = atomicAdd(head, 1);
        e = __shfl_sync(LL_FULL_MASK, e, 0);
        if (e >= n) break;
        const int4 en = P.queue[e < n_long ? e : P.queue_cap - 1 - (e - n_long)];
        const int i = en.y & ~(ASSOC_PLANE_BIT | ASSOC_OPEN_BIT);
        const bool open = (en.y & ASSOC_OPEN_BIT) != 0;
        const long long t0 = (P.dev_skip & 16) ? clock64() : 0;
        if (en.y & ASSOC_PLANE_BIT) assoc_warp_query<false>(P, en.x, i, open, en.z, en.w);
        else assoc_warp_query<true>(P, en.x, i, open, en.z, en.w);
        if ((P.dev_skip & 16) && lane_id() == 0) {  // development statistics: entries, cycles and longest entry per kind
            const int dt = (int)((clock64() - t0) >> 4), kind = open ? 0 : 1;
            atomicAdd(&P.lane[0].dbg[1 + kind], 1);
            atomicAdd(&P.lane[0].dbg[3 + kind], dt >> 6);
            atomicMax(&P.lane[0].dbg[5 + kind], dt);
        }
    }
}

// Few lanes (the single-stream path): a thread per query leaves most of the GPU idle and every query walks its buckets
// alone, so every query gets a WARP from the start - no thread pass, no queue; the same searches, the same total orders,
// hence the same correspondences.  Slot w = (lane, feature): corners first, then planes.
__global__ void __launch_bounds__(256, 4) k_odom_assoc_direct(OdomParams P, int n_lanes)
{
    const int nc = P.R * LL_SHARP_PER_RING, per_lane = nc + P.R * LL_FLAT_PER_RING, total = n_lanes * per_lane;
    for (int w = blockIdx.x * 8 + warp_id(); w < total; w += gridDim.x * 8) {
        const int b = w / per_lane, e = w - b * per_lane;
        const LaneState& L = P.lane[b];
        if (!L.inited || L.err) continue;
        if (e < nc) { if (e < L.n_sharp) assoc_warp_query<true>(P, b, e, true, (int)(unsigned)(NN_NONE >> 32), 0, true); }
        else if (e - nc < L.n_flat) assoc_warp_query<false>(P, b, e - nc, true, (int)(unsigned)(NN_NONE >> 32), 0, true);
    }
}

// ------------------------------------------------------------------------------------------------------
// compaction + graph vote + residual-block records; one CTA per lane
// ------------------------------------------------------------------------------------------------------
#define PREP_THREADS 1024
// Compaction of the matches in feature order and the residual-block records.  Corners: LidarEdgeFactor records
// (LO:556-618).  Planes: every match gets its LidarPlaneFactor_modify record (LF:210-211 normal) at slot
// ncorner + k with weight 1 (LO:781-787); when the graph vote is on (now_frame > 5) k_odom_vote then overwrites
// the weights and disables the rejected matches (record type -1), so the block order stays the reference's
// selected_idx order with gaps.
__global__ void __launch_bounds__(PREP_THREADS) k_odom_prep(OdomParams P)
{
    __shared__ int ws[40];
    const int maxp = P.R * LL_FLAT_PER_RING;
    const int b = blockIdx.x, tid = threadIdx.x;
    LaneState& L = P.lane[b];
    if (!L.inited || L.err) { if (tid == 0) L.n_blocks = 0; return; }
    const int ns = L.n_sharp, nf = L.n_flat, slot = last_slot(L);
    const float4* sharp = P.sharp + (size_t)b * P.R * LL_SHARP_PER_RING;
    const float4* flat = P.flat + (size_t)b * P.R * LL_FLAT_PER_RING;
    const float4* lastc = P.lsharp[slot] + (size_t)b * P.R * LL_LSHARP_PER_RING;
    const float4* lasts = P.lflat[slot] + (size_t)b * P.Nmax;
    const int* ca = P.corner_assoc + (size_t)b * P.R * LL_SHARP_PER_RING * 2;
    int* pa = P.plane_assoc + (size_t)b * P.R * LL_FLAT_PER_RING * 4;
    double* blk = P.blocks + (size_t)b * LL_BLOCK_DOUBLES * P.nblk_cap;
    float4* vsrc = P.vote_src + (size_t)b * maxp;
    float4* vtgt = P.vote_tgt + (size_t)b * maxp;
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
            if (P.distortion) blk[11 * cap + pos] = point_ratio(cp);   // feature_s, LO:569-573
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
        const int4 m = *reinterpret_cast<const int4*>(pa + i * 4);
        if (m.x >= 0) {
            const float4 cp = flat[i], pj = lasts[m.x], pl = lasts[m.y], pm = lasts[m.z];
            vsrc[pos] = make_float4(cp.x, cp.y, cp.z, __int_as_float(i));  // Corre_Match src / tgt, LO:753-754 (+ the feature index)
            vtgt[pos] = pj;
            // LF:210-211  ljm_norm = (j - l).cross(j - m); normalize()
            const double ax = (double)pj.x - (double)pl.x, ay = (double)pj.y - (double)pl.y, az = (double)pj.z - (double)pl.z;
            const double bx = (double)pj.x - (double)pm.x, by = (double)pj.y - (double)pm.y, bz = (double)pj.z - (double)pm.z;
            double nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
            const double z = nx * nx + ny * ny + nz * nz;
            if (z > 0.0) { const double nn = sqrt(z); nx = nx / nn; ny = ny / nn; nz = nz / nn; }
            const int o = ncorner + pos;
            blk[0 * cap + o] = 1.0;
            blk[1 * cap + o] = cp.x; blk[2 * cap + o] = cp.y; blk[3 * cap + o] = cp.z;
            blk[4 * cap + o] = pj.x; blk[5 * cap + o] = pj.y; blk[6 * cap + o] = pj.z;
            blk[7 * cap + o] = nx; blk[8 * cap + o] = ny; blk[9 * cap + o] = nz;
            blk[10 * cap + o] = 1.0;             // LO:783: weight 1 while now_frame <= 5
            if (P.distortion) blk[11 * cap + o] = point_ratio(cp);   // feature_s, LO:739-743
            pa[i * 4 + 3] = 1000;
            ++pos;
        }
    }
    if (tid == 0) {
        const bool vote = L.now_frame > P.graph_from_frame;  // LO:781 / LO:794
        L.n_blocks = ncorner + nplane;
        L.n_corner_corr = ncorner;
        L.n_plane_corr = nplane;
        L.corner_corr[P.outer] = ncorner;
        L.plane_corr[P.outer] = nplane;
        L.plane_sel[P.outer] = vote ? 0 : nplane;   // the vote kernels add the selected ones
    }
}

// graph_based_correspondence_vote_simple, plane case (LO:165-342): 10 contiguous regions, one CTA per (region, lane).
// Every unordered pair of the region is evaluated once (Distance() LO:153-162 is symmetric bit for bit).  The
// reference's test  expf(-(gap*gap)) < 0.96f  is  gap*gap >= t_min  with t_min calibrated on the host's glibc expf;
// gap = |sqrtf(d_src) - sqrtf(d_tgt)| is first estimated with the fast reciprocal square root, and only pairs whose
// estimate lies within 1e-3 of the threshold take the IEEE square roots that decide them exactly.
#define VOTE_THREADS 256
__global__ void __launch_bounds__(VOTE_THREADS) k_odom_vote(OdomParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int b = blockIdx.y, reg = blockIdx.x, tid = threadIdx.x;
    LaneState& L = P.lane[b];
    if (!L.inited || L.err || !(L.now_frame > P.graph_from_frame)) return;  // LO:781 / LO:794
    const int nplane = L.n_plane_corr, ncorner = L.n_corner_corr;
    const int region_len = nplane / 10;                              // LO:202-215
    const int r0 = region_len * reg, r1 = reg == 9 ? nplane : region_len * (reg + 1);
    const int m = r1 - r0;
    if (m <= 0) return;
    const int maxp = P.R * LL_FLAT_PER_RING;
    const int mcap = maxp / 10 + 16;
    float4* src = reinterpret_cast<float4*>(smem_raw);
    float4* tgt = src + mcap;
    int* votes = reinterpret_cast<int*>(tgt + mcap);
    __shared__ int nsel_s;
    for (int k = tid; k < m; k += VOTE_THREADS) {
        src[k] = P.vote_src[(size_t)b * maxp + r0 + k];
        tgt[k] = P.vote_tgt[(size_t)b * maxp + r0 + k];
        votes[k] = 0;
    }
    if (tid == 0) nsel_s = 0;
    __syncthreads();
    const float t_min = P.vote_t_min;
    const float g_mid = sqrtf(t_min);
    // rows k and m-1-k together hold m-1 pairs: every row pair is the same amount of work; four threads share one
    for (int k = tid >> 2; k < (m + 1) / 2; k += VOTE_THREADS / 4) {
        const int q = tid & 3;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const int row = half == 0 ? k : m - 1 - k;
            if (half == 1 && row == k) break;
            const float4 a = src[row], c = tgt[row];
            int mine = 0;
            for (int j = row + 1 + q; j < m; j += 4) {
                const float4 sj = src[j], tj = tgt[j];
                const float d1 = sqdist3(a.x, a.y, a.z, sj.x, sj.y, sj.z), d2 = sqdist3(c.x, c.y, c.z, tj.x, tj.y, tj.z);
                const float ge = fabsf(d1 * rsqrtf(fmaxf(d1, 1e-30f)) - d2 * rsqrtf(fmaxf(d2, 1e-30f)));
                bool v = ge > g_mid;
                if (fabsf(ge - g_mid) < 1e-3f) {  // too close to call with the approximation: LO:236-242 to the letter
                    const float gap = fabsf(sqrtf(d1) - sqrtf(d2));
                    v = gap * gap >= t_min;
                }
                if (v) { ++mine; atomicAdd(&votes[j], 1); }
            }
            if (mine) atomicAdd(&votes[row], mine);
        }
    }
    __syncthreads();
    double* blk = P.blocks + (size_t)b * LL_BLOCK_DOUBLES * P.nblk_cap;
    int* pa = P.plane_assoc + (size_t)b * P.R * LL_FLAT_PER_RING * 4;
    const float num_selected = 0.90f * (float)m;                    // LO:299-300
    int sel = 0;
    for (int k = tid; k < m; k += VOTE_THREADS) {
        const float fv = (float)votes[k];
        float w;
        if (fv > num_selected) w = 0.f;                              // LO:312-316: this and all worse are dropped
        else if (fv <= 50.f) w = 5.0f;                               // LO:317-318
        else w = 1.0f;
        const int o = ncorner + r0 + k;
        if (w > 0.f) { blk[10 * P.nblk_cap + o] = (double)w; ++sel; }
        else blk[o] = -1.0;                                          // not selected: no residual block (LO:797-808)
        pa[__float_as_int(src[k].w) * 4 + 3] = (int)(w * 1000.f);
    }
    if (sel) atomicAdd(&nsel_s, sel);
    __syncthreads();
    if (tid == 0 && nsel_s) atomicAdd(&L.plane_sel[P.outer], nsel_s);
}

// k_odom_prep + k_odom_vote in one launch (the default vote mode): one CTA per (region, lane).  Every CTA counts the lane's
// matches itself (two block scans over the association tables), so it knows the compacted order; it then writes the records of
// ITS share - a tenth of the corners and the plane matches of its vote region - and keeps the region's Corre_Match points in
// shared memory for the vote instead of sending them through global memory to a second kernel.  Same records, same votes.
__global__ void __launch_bounds__(VOTE_THREADS) k_odom_prep_vote(OdomParams P)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int ws[40];
    __shared__ int nsel_s;
    const int b = blockIdx.y, reg = blockIdx.x, tid = threadIdx.x;
    LaneState& L = P.lane[b];
    if (!L.inited || L.err) { if (reg == 0 && tid == 0) L.n_blocks = 0; return; }
    const int maxc = P.R * LL_SHARP_PER_RING, maxp = P.R * LL_FLAT_PER_RING;
    const int mcap = maxp / 10 + 16;
    float4* src = reinterpret_cast<float4*>(smem_raw);
    float4* tgt = src + mcap;
    int* votes = reinterpret_cast<int*>(tgt + mcap);
    const int ns = L.n_sharp, nf = L.n_flat, slot = last_slot(L);
    const bool vote = L.now_frame > P.graph_from_frame;   // LO:781 / LO:794
    const float4* sharp = P.sharp + (size_t)b * maxc;
    const float4* flat = P.flat + (size_t)b * maxp;
    const float4* lastc = P.lsharp[slot] + (size_t)b * P.R * LL_LSHARP_PER_RING;
    const float4* lasts = P.lflat[slot] + (size_t)b * P.Nmax;
    const int* ca = P.corner_assoc + (size_t)b * maxc * 2;
    int* pa = P.plane_assoc + (size_t)b * maxp * 4;
    double* blk = P.blocks + (size_t)b * LL_BLOCK_DOUBLES * P.nblk_cap;
    const int cap = P.nblk_cap;
    if (tid == 0) nsel_s = 0;

    // ---- corners (LO:556-618): compacted order from a scan over all of them, records for positions [c0, c1) ----------
    int ncorner = 0;
    {
        const int per = (maxc + VOTE_THREADS - 1) / VOTE_THREADS;
        const int i0 = min(tid * per, ns), i1 = min(i0 + per, ns);
        int mine = 0;
        for (int i = i0; i < i1; ++i) mine += ca[i * 2 + 1] >= 0;
        int pos = block_exclusive_scan(mine, ws, &ncorner);
        const int c0 = (int)((long long)ncorner * reg / 10), c1 = (int)((long long)ncorner * (reg + 1) / 10);
        for (int i = i0; i < i1; ++i) {
            const int2 m = *reinterpret_cast<const int2*>(ca + i * 2);
            if (m.y < 0) continue;
            if (pos >= c0 && pos < c1) {
                const float4 cp = sharp[i], a = lastc[m.x], c = lastc[m.y];
                blk[0 * cap + pos] = 0.0;
                blk[1 * cap + pos] = cp.x; blk[2 * cap + pos] = cp.y; blk[3 * cap + pos] = cp.z;
                blk[4 * cap + pos] = a.x; blk[5 * cap + pos] = a.y; blk[6 * cap + pos] = a.z;
                blk[7 * cap + pos] = c.x; blk[8 * cap + pos] = c.y; blk[9 * cap + pos] = c.z;
                blk[10 * cap + pos] = 1.0;
                if (P.distortion) blk[11 * cap + pos] = point_ratio(cp);   // feature_s, LO:569-573
            }
            ++pos;
        }
    }
    // ---- planes: compacted order, this CTA's vote region [r0, r1) (LO:202-215) ----------------------------------------
    int nplane = 0;
    const int per = (maxp + VOTE_THREADS - 1) / VOTE_THREADS;
    const int i0 = min(tid * per, nf), i1 = min(i0 + per, nf);
    int mine = 0;
    for (int i = i0; i < i1; ++i) mine += pa[i * 4] >= 0;
    int pos = block_exclusive_scan(mine, ws, &nplane);
    const int region_len = nplane / 10;
    const int r0 = region_len * reg, r1 = reg == 9 ? nplane : region_len * (reg + 1);
    const int m = r1 - r0;
    for (int i = i0; i < i1; ++i) {
        const int4 mt = *reinterpret_cast<const int4*>(pa + i * 4);
        if (mt.x < 0) continue;
        if (pos >= r0 && pos < r1) {
            const float4 cp = flat[i], pj = lasts[mt.x], pl = lasts[mt.y], pm = lasts[mt.z];
            src[pos - r0] = make_float4(cp.x, cp.y, cp.z, __int_as_float(i));  // Corre_Match src / tgt, LO:753-754 (+ the feature index)
            tgt[pos - r0] = pj;
            votes[pos - r0] = 0;
            // LF:210-211  ljm_norm = (j - l).cross(j - m); normalize()
            const double ax = (double)pj.x - (double)pl.x, ay = (double)pj.y - (double)pl.y, az = (double)pj.z - (double)pl.z;
            const double bx = (double)pj.x - (double)pm.x, by = (double)pj.y - (double)pm.y, bz = (double)pj.z - (double)pm.z;
            double nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
            const double z = nx * nx + ny * ny + nz * nz;
            if (z > 0.0) { const double nn = sqrt(z); nx = nx / nn; ny = ny / nn; nz = nz / nn; }
            const int o = ncorner + pos;
            blk[0 * cap + o] = 1.0;
            blk[1 * cap + o] = cp.x; blk[2 * cap + o] = cp.y; blk[3 * cap + o] = cp.z;
            blk[4 * cap + o] = pj.x; blk[5 * cap + o] = pj.y; blk[6 * cap + o] = pj.z;
            blk[7 * cap + o] = nx; blk[8 * cap + o] = ny; blk[9 * cap + o] = nz;
            blk[10 * cap + o] = 1.0;             // LO:783: weight 1 while now_frame <= 5
            if (P.distortion) blk[11 * cap + o] = point_ratio(cp);   // feature_s, LO:739-743
            if (!vote) pa[i * 4 + 3] = 1000;
        }
        ++pos;
    }
    if (reg == 0 && tid == 0) {
        L.n_blocks = ncorner + nplane;
        L.n_corner_corr = ncorner;
        L.n_plane_corr = nplane;
        L.corner_corr[P.outer] = ncorner;
        L.plane_corr[P.outer] = nplane;
    }
    if (!vote) {   // every match is selected with weight 1 (LO:781-787)
        if (tid == 0 && m > 0) atomicAdd(&L.plane_sel[P.outer], m);
        return;
    }
    __syncthreads();
    if (m <= 0) return;
    // ---- graph_based_correspondence_vote_simple on the region (LO:165-342), as k_odom_vote ---------------------------------
    const float t_min = P.vote_t_min;
    const float g_mid = sqrtf(t_min);
    for (int k = tid >> 2; k < (m + 1) / 2; k += VOTE_THREADS / 4) {
        const int q = tid & 3;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            const int row = half == 0 ? k : m - 1 - k;
            if (half == 1 && row == k) break;
            const float4 a = src[row], c = tgt[row];
            int cnt = 0;
            for (int j = row + 1 + q; j < m; j += 4) {
                const float4 sj = src[j], tj = tgt[j];
                const float d1 = sqdist3(a.x, a.y, a.z, sj.x, sj.y, sj.z), d2 = sqdist3(c.x, c.y, c.z, tj.x, tj.y, tj.z);
                const float ge = fabsf(d1 * rsqrtf(fmaxf(d1, 1e-30f)) - d2 * rsqrtf(fmaxf(d2, 1e-30f)));
                bool v = ge > g_mid;
                if (fabsf(ge - g_mid) < 1e-3f) {  // too close to call with the approximation: LO:236-242 to the letter
                    const float gap = fabsf(sqrtf(d1) - sqrtf(d2));
                    v = gap * gap >= t_min;
                }
                if (v) { ++cnt; atomicAdd(&votes[j], 1); }
            }
            if (cnt) atomicAdd(&votes[row], cnt);
        }
    }
    __syncthreads();
    const float num_selected = 0.90f * (float)m;                    // LO:299-300
    int sel = 0;
    for (int k = tid; k < m; k += VOTE_THREADS) {
        const float fv = (float)votes[k];
        float w;
        if (fv > num_selected) w = 0.f;                              // LO:312-316: this and all worse are dropped
        else if (fv <= 50.f) w = 5.0f;                               // LO:317-318
        else w = 1.0f;
        const int o = ncorner + r0 + k;
        if (w > 0.f) { blk[10 * cap + o] = (double)w; ++sel; }
        else blk[o] = -1.0;                                          // not selected: no residual block (LO:797-808)
        pa[__float_as_int(src[k].w) * 4 + 3] = (int)(w * 1000.f);
    }
    if (sel) atomicAdd(&nsel_s, sel);
    __syncthreads();
    if (tid == 0 && nsel_s) atomicAdd(&L.plane_sel[P.outer], nsel_s);
}

// graph_based_correspondence_vote_partial (laserMapping.cpp:321-834, graph_construction_partial LM:261-318): the paper-style
// scoring, DEAD in the reference (its only call is commented out, LO:622).  Optional mode cfg.vote_mode = 1, beyond-reference:
// it replaces vote_simple on the odometry's plane correspondences.  One CTA per (region, lane); the region's m x m
// compatibility matrix G(i,j) = expf(-gap^2) lives in shared memory (expf and cbrt evaluated in fp64 and rounded once,
// which is what glibc's float routines return in all but double-rounding cases).  Steps as in the source: neighbours
// G > 0.95; first score = mean over neighbour pairs of cbrt(G_ia G_ib G_ab); threshold = min(ratio of the sums, mean
// score); pruning; final score = 0.1 * mean G(a,i) + 0.9 * (pairs with G_ab != 0) / (d (d - 2) / 2) - the source's
// pow(x, 1 / 3) has the integer exponent 0 - for d > 2; selected = score != 0 with weight = score.
#define VP_THREADS 256
#define VP_WORDS 8   // neighbour bitmask words per vertex: regions of up to 256 correspondences
__global__ void __launch_bounds__(VP_THREADS) k_odom_vote_partial(OdomParams P, int mcap)
{
    extern __shared__ __align__(16) unsigned char vp_smem[];
    const int b = blockIdx.y, reg = blockIdx.x, tid = threadIdx.x, lane = lane_id(), wid = warp_id();
    LaneState& L = P.lane[b];
    if (!L.inited || L.err || !(L.now_frame > P.graph_from_frame)) return;
    const int nplane = L.n_plane_corr, ncorner = L.n_corner_corr;
    const int region_len = nplane / 10;
    const int r0 = region_len * reg, r1 = reg == 9 ? nplane : region_len * (reg + 1);
    const int m = r1 - r0;
    if (m <= 0) return;
    const int maxp = P.R * LL_FLAT_PER_RING;
    float* G = reinterpret_cast<float*>(vp_smem);                       // [m][m]
    unsigned* conn = reinterpret_cast<unsigned*>(G + (size_t)mcap * mcap);   // [m][VP_WORDS]
    float* score = reinterpret_cast<float*>(conn + (size_t)mcap * VP_WORDS);  // [m] first-pass score
    float* numer = score + mcap;                                        // [m]
    float* fin = numer + mcap;                                          // [m] final score
    __shared__ float thr_s;
    __shared__ int any_s, nsel_s;
    const float4* src = P.vote_src + (size_t)b * maxp + r0;
    const float4* tgt = P.vote_tgt + (size_t)b * maxp + r0;
    double* blk = P.blocks + (size_t)b * LL_BLOCK_DOUBLES * P.nblk_cap;
    int* pa = P.plane_assoc + (size_t)b * P.R * LL_FLAT_PER_RING * 4;
    if (tid == 0) { any_s = 0; nsel_s = 0; }
    if (m > mcap || m > 32 * VP_WORDS) {   // cannot happen with the configured capacities; never leave the blocks half-decided
        for (int k = tid; k < m; k += VP_THREADS) blk[ncorner + r0 + k] = -1.0;
        return;
    }
    for (int k = tid; k < m * VP_WORDS; k += VP_THREADS) conn[k] = 0u;
    __syncthreads();
    // graph_construction_partial
    int any = 0;
    for (int p = tid; p < m * m; p += VP_THREADS) {
        const int i = p / m, j = p % m;
        float g = 0.f;
        if (i != j) {
            const float4 a = __ldg(src + i), c = __ldg(src + j), ta = __ldg(tgt + i), tc = __ldg(tgt + j);
            const float s1 = sqrtf(sqdist3(a.x, a.y, a.z, c.x, c.y, c.z)), s2 = sqrtf(sqdist3(ta.x, ta.y, ta.z, tc.x, tc.y, tc.z));
            const float gap = fabsf(s1 - s2);
            g = (float)exp(-(double)(gap * gap));
            if ((double)g > 0.95) atomicOr(&conn[i * VP_WORDS + (j >> 5)], 1u << (j & 31));
            any |= g != 0.f;
        }
        G[i * m + j] = g;
    }
    if (any) any_s = 1;
    __syncthreads();
    if (!any_s) {   // LM:399-403: "Graph is not connected!" -> nothing selected from this region
        for (int k = tid; k < m; k += VP_THREADS) { blk[ncorner + r0 + k] = -1.0; pa[__float_as_int(__ldg(src + k).w) * 4 + 3] = 0; }
        return;
    }
    auto nth_words = [&](const unsigned* w, int words) { int d = 0; for (int q = 0; q < words; ++q) d += __popc(w[q]); return d; };
    const int words = (m + 31) >> 5;
    // first pass (LM:445-500): one warp per vertex, lanes share the first neighbour of the pair
    for (int i = wid; i < m; i += VP_THREADS / 32) {
        const unsigned* ci = conn + i * VP_WORDS;
        const int deg = nth_words(ci, words);
        double acc = 0.0;
        for (int a = lane; a < m; a += 32) {
            if (!((ci[a >> 5] >> (a & 31)) & 1u)) continue;
            const float gia = G[i * m + a];
            for (int bq = a + 1; bq < m; ++bq) {
                if (!((ci[bq >> 5] >> (bq & 31)) & 1u)) continue;
                const float gab = G[a * m + bq];
                if (gab != 0.f) acc += (double)(float)pow((double)(gia * G[i * m + bq] * gab), 1.0 / 3);
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) acc += shfl_xor_f64(acc, d);
        if (lane == 0) {
            float sc = 0.f, nu = 0.f;
            if (deg > 1) { nu = (float)acc; sc = nu / (float)(deg * (deg - 1) * 0.5); }
            score[i] = sc; numer[i] = deg > 1 ? nu : -1.f;   // -1: does not enter the filter sums
            fin[i] = (float)deg;                              // parked for thread 0
        }
    }
    __syncthreads();
    if (tid == 0) {   // LM:493-503, vertex order, fp32
        float fa_n = 0.f, fa_d = 0.f, fb = 0.f;
        for (int i = 0; i < m; ++i) {
            if (numer[i] >= 0.f) { const int deg = (int)fin[i]; fa_n += numer[i]; fa_d += (float)(deg * (deg - 1) * 0.5); }
            fb += score[i];
        }
        const float fa = fa_n / fa_d;
        fb = fb / (float)m;
        thr_s = fminf(fa, fb);   // std::min: a NaN ratio (no vertex with two neighbours) never wins over fb... std::min(a, b) = b < a ? b : a
        if (!(fb < fa)) thr_s = fa;
    }
    __syncthreads();
    const float thr = thr_s;
    // prune (LM:560-580): a neighbour stays when its first-pass score reaches the threshold
    for (int k = tid; k < m * words; k += VP_THREADS) {
        const int i = k / words, q = k % words;
        unsigned w = conn[i * VP_WORDS + q], keep = 0u;
        while (w) { const int bit = __ffs(w) - 1; w &= w - 1; if (score[q * 32 + bit] >= thr) keep |= 1u << bit; }
        conn[i * VP_WORDS + q] = keep;
    }
    __syncthreads();
    // final score (LM:600-690)
    for (int i = wid; i < m; i += VP_THREADS / 32) {
        const unsigned* ci = conn + i * VP_WORDS;
        const int d = nth_words(ci, words);
        double loose = 0.0;
        int tight = 0;
        if (d > 2) {
            for (int a = lane; a < m; a += 32) {
                if (!((ci[a >> 5] >> (a & 31)) & 1u)) continue;
                loose += (double)G[a * m + i];
                for (int bq = a + 1; bq < m; ++bq)
                    if (((ci[bq >> 5] >> (bq & 31)) & 1u) && G[a * m + bq] != 0.f) ++tight;
            }
        }
#pragma unroll
        for (int dd = 16; dd > 0; dd >>= 1) { loose += shfl_xor_f64(loose, dd); tight += __shfl_xor_sync(LL_FULL_MASK, tight, dd); }
        if (lane == 0) {
            float tight_sum = 0.f, looser_sum = 0.f;
            if (d > 2) { tight_sum = (float)(double)tight; tight_sum /= (float)(d * (d - 2) / 2); }
            if (d != 0) { looser_sum = (float)loose; looser_sum = looser_sum / (float)d; }
            const float wb = 0.9f;
            fin[i] = (1 - wb) * looser_sum + wb * tight_sum;
        }
    }
    __syncthreads();
    int sel = 0;
    for (int k = tid; k < m; k += VP_THREADS) {
        const float w = fin[k];
        const int o = ncorner + r0 + k;
        if (w != 0.f) { blk[10 * P.nblk_cap + o] = (double)w; ++sel; }
        else blk[o] = -1.0;
        pa[__float_as_int(__ldg(src + k).w) * 4 + 3] = (int)(w * 1000.f);
    }
    if (sel) atomicAdd(&nsel_s, sel);
    __syncthreads();
    if (tid == 0 && nsel_s) atomicAdd(&L.plane_sel[P.outer], nsel_s);
}

// grid (lanes, parts): with few lanes `parts` CTAs share one lane's residual blocks and all-reduce the 28 doubles of every
// evaluation through the context's mailbox (LmComm, gworld = 1) - the single-stream latency path; parts = 1 is the plain solve
template <bool DIST>
__global__ void __launch_bounds__(LM_THREADS) k_lm_solve_odom(OdomParams P, LmComm comm, int n_all_lanes)
{
    const int b = blockIdx.x, part = blockIdx.y;
    LaneState& L = P.lane[b];
    const bool split = comm.nparts > 1, clus = comm.cluster > 1;
    if (clus) {   // the CTAs of a cluster work on one lane: they take the same decision here
        if (!L.inited || L.err) return;
        lm_solve<DIST>(P.blocks + (size_t)b * LL_BLOCK_DOUBLES * P.nblk_cap, P.nblk_cap, L.n_blocks, L.para_q, L.para_t, &L, P.outer, &comm, b, part);
        return;
    }
    if (split && b == 0 && part == 0)   // lanes outside this launch keep their collective counters
        for (int i = gridDim.x + threadIdx.x; i < n_all_lanes; i += blockDim.x) comm.seq_out[i] = comm.seq_in[i];
    if (!L.inited || L.err) {
        if (split && part == 0 && threadIdx.x == 0) comm.seq_out[b] = comm.seq_in[b];
        return;
    }
    lm_solve<DIST>(P.blocks + (size_t)b * LL_BLOCK_DOUBLES * P.nblk_cap, P.nblk_cap, L.n_blocks, L.para_q, L.para_t, &L, P.outer, split ? &comm : nullptr, b, part);
}

// LO:861-880 TransformToEnd of this frame's less-sharp / less-flat clouds (they become *Last at the swap), distortion == 2:
// the reference keeps the block behind a literal `if (0)`; with it enabled the clouds are expressed at the scan's end and
// lose their time fraction (intensity = int(intensity)).  Frames that only initialise are left alone.
__global__ void k_odom_to_end(OdomParams P, float4* ls0, float4* ls1, float4* lf0, float4* lf1)
{
    const int b = blockIdx.y;
    const LaneState& L = P.lane[b];
    if (!L.inited || L.err) return;
    float4* ls = (L.cur == 0 ? ls0 : ls1) + (size_t)b * P.R * LL_LSHARP_PER_RING;
    float4* lf = (L.cur == 0 ? lf0 : lf1) + (size_t)b * P.Nmax;
    const int nc = L.n_less_sharp, ns = L.n_less_flat;
    const double n2 = L.para_q[0] * L.para_q[0] + L.para_q[1] * L.para_q[1] + L.para_q[2] * L.para_q[2] + L.para_q[3] * L.para_q[3];
    double qi[4] = {0, 0, 0, 0};   // Eigen inverse(): conjugate / squaredNorm
    if (n2 > 0.0) { qi[0] = -L.para_q[0] / n2; qi[1] = -L.para_q[1] / n2; qi[2] = -L.para_q[2] / n2; qi[3] = L.para_q[3] / n2; }
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nc + ns; e += gridDim.x * blockDim.x) {
        float4* pp = e < nc ? ls + e : lf + (e - nc);
        const float4 p = *pp;
        const double s = point_ratio(p);
        double qs[4], sx, sy, sz, ex, ey, ez;
        lm_identity_slerp<double>(s, L.para_q, qs);
        quat_rotate(qs, (double)p.x, (double)p.y, (double)p.z, sx, sy, sz);
        const float ux = (float)(sx + s * L.para_t[0]), uy = (float)(sy + s * L.para_t[1]), uz = (float)(sz + s * L.para_t[2]);   // un_point_tmp is a PointXYZI: fp32
        quat_rotate(qi, (double)ux - L.para_t[0], (double)uy - L.para_t[1], (double)uz - L.para_t[2], ex, ey, ez);
        *pp = make_float4((float)ex, (float)ey, (float)ez, (float)(int)p.w);
    }
}

// LO:830-831 pose accumulation; LO:882-896 swap (the grids are rebuilt right after); counters LO:925-926
__global__ void k_odom_finalize(LaneState* lane, double* pose_out, int* status, int n_lanes)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_lanes) return;
    LaneState& L = lane[b];
    status[b] = L.err;
    if (L.err) {
        // a scan the feature stage rejected (no valid point, ring over capacity) does not advance the stream: pose, *Last
        // clouds, warm start and frame counter stay as they were; the caller sees the code in the lane's status
        if (pose_out) {
            double* o = pose_out + (size_t)b * 14;
            for (int k = 0; k < 4; ++k) { o[k] = L.q_w[k]; o[7 + k] = L.q_w[k]; }
            for (int k = 0; k < 3; ++k) { o[4 + k] = L.t_w[k]; o[11 + k] = L.t_w[k]; }
        }
        return;
    }
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

static int build_grid(ll_ctx* c, KnnGrid& g, const float4* p0, const float4* p1, size_t lane_stride, int which, int n_lanes, int max_pts)
{
    GridSrc S;
    S.pts[0] = p0; S.pts[1] = p1; S.lane_stride = lane_stride; S.which = which;
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

#ifdef LL_ASSOC_STATS
extern "C" void ll_dev_assoc_stats(unsigned long long out[32], int reset)
{
    cudaMemcpyFromSymbol(out, g_assoc_stats, sizeof(unsigned long long) * 32);
    if (reset) { unsigned long long z[32] = {0}; cudaMemcpyToSymbol(g_assoc_stats, z, sizeof(z)); }
}
#endif

int ll_launch_odometry(ll_ctx* c, int n_lanes)
{
    OdomParams P;
    P.lane = c->d_lane; P.sharp = c->d_sharp; P.flat = c->d_flat;
    P.lsharp[0] = c->d_lsharp[0]; P.lsharp[1] = c->d_lsharp[1]; P.lflat[0] = c->d_lflat[0]; P.lflat[1] = c->d_lflat[1];
    P.Nmax = c->Nmax; P.R = c->R; P.ac = c->a_corner; P.as_ = c->a_surf; P.bands[0] = c->d_bands[0]; P.bands[1] = c->d_bands[1]; P.az_bins_corner = c->az_bins_corner; P.az_bins_surf = c->az_bins_surf;
    P.corner_assoc = c->d_corner_assoc; P.plane_assoc = c->d_plane_assoc; P.blocks = c->d_blocks; P.nblk_cap = c->nblk_cap;
    P.distortion = c->cfg.distortion; P.graph_from_frame = c->cfg.graph_from_frame; P.vote_t_min = c->vote_t_min; P.dev_skip = getenv("LL_DEV_SKIP") ? atoi(getenv("LL_DEV_SKIP")) : 0;
    P.vote_src = c->d_vote_src; P.vote_tgt = c->d_vote_tgt;
    P.queue = c->d_assoc_queue; P.queue_n = c->d_assoc_queue_n; P.queue_cap = c->assoc_queue_cap;
    cudaStream_t s = c->stream;
    // threads per CTA of the association's thread pass: smaller CTAs retire sooner (the queries' costs differ a lot)
    int assoc_threads = getenv("LL_ASSOC_THREADS") ? atoi(getenv("LL_ASSOC_THREADS")) : ASSOC_THREADS;
    if (assoc_threads != 32 && assoc_threads != 64 && assoc_threads != 128) assoc_threads = ASSOC_THREADS;
    if (assoc_threads < c->R) assoc_threads = c->R <= 64 ? 64 : 128;   // the CTA's first R threads stage the ring bands
    const int cblocks = (c->R * LL_SHARP_PER_RING + assoc_threads - 1) / assoc_threads, pblocks = (c->R * LL_FLAT_PER_RING + assoc_threads - 1) / assoc_threads;
    LL_CUDA_CHECK(c, cudaMemsetAsync(c->d_assoc_queue_n, 0, sizeof(int) * 16, s));
    const int heavy_blocks = getenv("LL_HEAVY_BLOCKS") ? atoi(getenv("LL_HEAVY_BLOCKS")) : 148 * 4;
    const int dmax = getenv("LL_ASSOC_DMAX") ? atoi(getenv("LL_ASSOC_DMAX")) : 8;   // ring-window bins per side a thread walks
    const int minb = getenv("LL_ASSOC_MINB") ? atoi(getenv("LL_ASSOC_MINB")) : 8;   // resident blocks per SM the thread pass is compiled for
    // The solve needs 128 registers per thread: 512-thread CTAs fill an SM's register file alone.  With more problems than
    // SMs, 256-thread CTAs run two per SM in ONE wave and fill each other's serial phases (LM controller, barriers).
    int n_sm = 148;
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, c->dev);
    const int lm_threads = getenv("LL_LM_THREADS") ? atoi(getenv("LL_LM_THREADS")) : (n_lanes > n_sm ? 256 : LM_THREADS);
    // Splitting a solve over several CTAs through the MAILBOX (LL_LM_PARTS = 2..16, lanes x parts <= SMs; the form the
    // multi-GPU scan-to-map solve uses) does not pay here: with ~1900 blocks the global-memory all-reduce of every evaluation
    // costs more than the split saves (single stream on B200: 0.124 ms per solve with 16 parts against 0.074 ms with one
    // CTA), so its default is 1.  The split that does pay on one GPU is the thread-block cluster below.
    int lm_parts = 1;
    if (const char* e = getenv("LL_LM_PARTS")) { const int v = atoi(e); if (v >= 1 && v <= LM_MAX_PARTS && v * n_lanes <= n_sm) lm_parts = v; }
    LmComm comm;
    for (int g = 0; g < LM_MAX_GPUS; ++g) { comm.mbox[g] = nullptr; comm.flag[g] = nullptr; }
    comm.mbox[0] = reinterpret_cast<double*>(c->d_odom_comm);
    comm.flag[0] = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(c->d_odom_comm) + c->odom_comm_mbox_bytes);
    comm.grank = 0; comm.gworld = 1; comm.nparts = lm_parts; comm.timeout_ns = 2000000000ull;
    const int lm_threads_used = lm_parts > 1 ? (getenv("LL_LM_THREADS") ? lm_threads : 128) : lm_threads;   // split: ~1 block per thread
    // Few lanes (the single-stream path): the CTAs of a thread-block cluster share a lane's blocks and sum their 28 doubles
    // through distributed shared memory - one hardware cluster barrier per evaluation, no mailbox, no cooperative launch
    int lm_cluster = lm_parts > 1 ? 1 : lm_cluster_size(n_lanes, n_sm);
    if (lm_cluster > 1) {
        int& ok = c->cluster_ok[P.distortion ? 1 : 0];
        if (ok < 0) ok = (P.distortion ? lm_cluster_fits(k_lm_solve_odom<true>, 8, LM_THREADS) : lm_cluster_fits(k_lm_solve_odom<false>, 8, LM_THREADS)) ? 1 : 0;
        if (!ok) lm_cluster = 1;
    }
    comm.cluster = lm_cluster;
    int lm_cluster_threads = ((c->R * (LL_SHARP_PER_RING + LL_FLAT_PER_RING) + lm_cluster - 1) / lm_cluster + 63) / 64 * 64;   // ~1 block per thread
    lm_cluster_threads = lm_cluster_threads < 64 ? 64 : (lm_cluster_threads > LM_THREADS ? LM_THREADS : lm_cluster_threads);
    if (getenv("LL_LM_THREADS")) lm_cluster_threads = lm_threads;
    const int kmax = getenv("LL_ASSOC_KMAX") ? atoi(getenv("LL_ASSOC_KMAX")) : 2;   // 1-NN bins per side a thread walks
    const size_t vote_smem = (size_t)(c->R * LL_FLAT_PER_RING / 10 + 16) * (2 * sizeof(float4) + sizeof(int));
    // kdtreeCornerLast / kdtreeSurfLast ->setInputCloud (LO:895-896), deferred to the moment the trees are queried:
    // the polar indexes of the two *Last clouds are built together (4 launches) right before the association, so the
    // tables are still in L2 when the queries walk them
    {
        IndexSet S;
        KnnGrid* tab[2] = {&c->a_corner, &c->a_surf};
        S.pts[0][0] = c->d_lsharp[0]; S.pts[0][1] = c->d_lsharp[1]; S.pts[1][0] = c->d_lflat[0]; S.pts[1][1] = c->d_lflat[1];
        S.lane_stride[0] = (size_t)c->R * LL_LSHARP_PER_RING; S.lane_stride[1] = (size_t)c->Nmax;
        S.chunk_begin[0] = 0;
        for (int t = 0; t < 2; ++t) {
            S.cursor[t] = tab[t]->cursor; S.start[t] = tab[t]->start; S.partial[t] = tab[t]->partial; S.sorted[t] = tab[t]->sorted;
            S.ebound[t] = c->d_ebound[t]; S.bands[t] = c->d_bands[t];
            S.T[t] = tab[t]->T; S.cap[t] = tab[t]->cap;
            S.chunk_begin[t + 1] = S.chunk_begin[t] + tab[t]->T / GRID_CHUNK;
            LL_CUDA_CHECK(c, cudaMemsetAsync(tab[t]->cursor, 0, sizeof(int) * (size_t)tab[t]->T * n_lanes, s));
            LL_CUDA_CHECK(c, cudaMemsetAsync(c->d_ebound[t], 0, sizeof(unsigned) * (size_t)c->R * 2 * n_lanes, s));
        }
        S.az_bins[0] = c->az_bins_corner; S.az_bins[1] = c->az_bins_surf; S.rings = c->R;
        // blocks of 256 threads x IDX_UNROLL points: the corner cloud holds at most R * 120 points, the surf cloud usually
        // about a quarter of the scan (more is covered by the grid-stride loops)
        const int per_block = 256 * IDX_UNROLL;
        S.gx_corner = (c->R * LL_LSHARP_PER_RING + per_block - 1) / per_block;
        const int gx_surf = (c->Nmax * 5 / 16 + per_block - 1) / per_block;
        const dim3 gidx(S.gx_corner + gx_surf, n_lanes);
        { LLProf pr(c, "k_index_count"); k_index_count<<<gidx, 256, 0, s>>>(S, c->d_lane); }
        { LLProf pr(c, "k_index_partial"); k_index_partial<<<dim3(S.chunk_begin[2], n_lanes), 256, 0, s>>>(S, c->d_lane); }
        { LLProf pr(c, "k_index_scan"); k_index_scan<<<dim3(S.chunk_begin[2], n_lanes), 256, 0, s>>>(S); }
        { LLProf pr(c, "k_index_scatter"); k_index_scatter<<<gidx, 256, 0, s>>>(S, c->d_lane); }
        c->launches += 4;
    }
    // slab form of the thread pass (LL_ASSOC_SLAB=1) or the global-memory form
    const bool slab_mode = getenv("LL_ASSOC_SLAB") && atoi(getenv("LL_ASSOC_SLAB")) == 1;
    SlabParams Sp;
    auto env_pow2 = [](const char* name, int dflt, int lo, int hi) { const char* e = getenv(name); int v = e ? atoi(e) : dflt; if (v < lo || v > hi || (v & (v - 1))) v = dflt; return v; };
    auto env_int = [](const char* name, int dflt, int lo, int hi) { const char* e = getenv(name); const int v = e ? atoi(e) : dflt; return v < lo || v > hi ? dflt : v; };
    Sp.qa = c->d_qa; Sp.qb = c->d_qb; Sp.qstart = c->d_qstart;
    Sp.slab_surf = env_pow2("LL_SLAB_SURF", 16, 1, c->az_bins_surf); Sp.slab_corner = env_pow2("LL_SLAB_CORNER", 8, 1, c->az_bins_corner);
    while (c->az_bins_surf / Sp.slab_surf > 128) Sp.slab_surf <<= 1;
    while (c->az_bins_corner / Sp.slab_corner > 128) Sp.slab_corner <<= 1;
    Sp.ns_surf = c->az_bins_surf / Sp.slab_surf; Sp.ns_corner = c->az_bins_corner / Sp.slab_corner;
    Sp.halo_surf = env_int("LL_HALO_SURF", 4, 1, 64); Sp.halo_corner = env_int("LL_HALO_CORNER", 2, 1, 64);
    const int slab_minb = env_int("LL_SLAB_MINB", 3, 1, 3);
    const size_t hdr_bins = (size_t)((Sp.slab_surf + 2 * Sp.halo_surf) > (Sp.slab_corner + 2 * Sp.halo_corner) ? (Sp.slab_surf + 2 * Sp.halo_surf) : (Sp.slab_corner + 2 * Sp.halo_corner));
    const size_t hdr_bytes = ((hdr_bins * c->R + 1) * 4 + 15) / 16 * 16;
    Sp.pcap = env_int("LL_SLAB_PCAP", 4032, 64, 13000);
    const size_t slab_smem = hdr_bytes + (size_t)Sp.pcap * 16;
    if (slab_mode && !c->slab_attr_set) {
        LL_CUDA_CHECK(c, cudaFuncSetAttribute(k_odom_assoc_slab<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        LL_CUDA_CHECK(c, cudaFuncSetAttribute(k_odom_assoc_slab<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024));
        LL_CUDA_CHECK(c, cudaFuncSetAttribute(k_odom_assoc_slab<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024));
        c->slab_attr_set = true;
    }
    if (slab_mode && slab_smem > (size_t)(slab_minb >= 3 ? 72 : (slab_minb == 2 ? 110 : 220)) * 1024) { c->last_error = "slab stage larger than the shared memory of the chosen occupancy"; return LL_E_INVAL; }
    if (Sp.ns_surf + Sp.ns_corner + 2 > c->qstart_stride) { c->last_error = "too many slabs"; return LL_E_INVAL; }
    // direct form (a warp per query, one launch) up to LL_ASSOC_DIRECT lanes
    const int direct_lanes = getenv("LL_ASSOC_DIRECT") ? atoi(getenv("LL_ASSOC_DIRECT")) : 8;
    const bool direct_mode = !slab_mode && n_lanes <= direct_lanes;
    for (int outer = 0; outer < 3; ++outer) {  // LO:439
        P.outer = outer;
        if (direct_mode) {
            LLProf pr(c, "k_odom_assoc_direct");
            const int want = (n_lanes * c->R * (LL_SHARP_PER_RING + LL_FLAT_PER_RING) + 7) / 8;
            k_odom_assoc_direct<<<want < n_sm * 16 ? want : n_sm * 16, 256, 0, s>>>(P, n_lanes);
            c->launches -= 1;   // one launch instead of the thread pass + the queue pass (5 counted per iteration below)
        } else if (slab_mode) {
            { LLProf pr(c, "k_odom_queries"); k_odom_queries<<<n_lanes, QPREP_THREADS, 0, s>>>(P, Sp); }
            LLProf pr(c, "k_odom_assoc");
            const dim3 g(Sp.ns_surf + Sp.ns_corner, n_lanes);
            if (slab_minb >= 3) k_odom_assoc_slab<3><<<g, SLAB_THREADS, slab_smem, s>>>(P, Sp);
            else if (slab_minb == 2) k_odom_assoc_slab<2><<<g, SLAB_THREADS, slab_smem, s>>>(P, Sp);
            else k_odom_assoc_slab<1><<<g, SLAB_THREADS, slab_smem, s>>>(P, Sp);
            c->launches += 1;
        } else {
            LLProf pr(c, "k_odom_assoc");
            const dim3 g(cblocks + pblocks, n_lanes);
            if (minb >= 12) k_odom_assoc<12><<<g, assoc_threads, 0, s>>>(P, cblocks, dmax, kmax);
            else if (minb >= 10) k_odom_assoc<10><<<g, assoc_threads, 0, s>>>(P, cblocks, dmax, kmax);
            else if (minb >= 8) k_odom_assoc<8><<<g, assoc_threads, 0, s>>>(P, cblocks, dmax, kmax);
            else k_odom_assoc<6><<<g, assoc_threads, 0, s>>>(P, cblocks, dmax, kmax);
        }
        if (!direct_mode) { LLProf pr(c, "k_odom_assoc_heavy"); k_odom_assoc_heavy<<<heavy_blocks, 256, 0, s>>>(P); }
        // records + vote in one launch for few lanes (two dependent launches less per iteration: 0.440 -> 0.423 ms per scan on one
        // stream); with many lanes the ten CTAs per lane each repeating the match count cost more than the launch saves
        // (3.61 vs 3.56 ms per 256-lane step), so the batched path keeps k_odom_prep + k_odom_vote.  LL_VOTE_FUSED = 0 / 1 forces.
        const char* fv_env = getenv("LL_VOTE_FUSED");
        const bool fused_vote = c->cfg.vote_mode != 1 && (fv_env ? atoi(fv_env) != 0 : n_lanes <= 16);
        if (fused_vote) {
            LLProf pr(c, "k_odom_prep_vote"); k_odom_prep_vote<<<dim3(10, n_lanes), VOTE_THREADS, vote_smem + 64, s>>>(P);
            c->launches -= 1;   // one launch for the records and the vote (5 counted per iteration below)
        } else { LLProf pr(c, "k_odom_prep"); k_odom_prep<<<n_lanes, PREP_THREADS, 0, s>>>(P); }
        if (fused_vote) {
        } else if (c->cfg.vote_mode == 1) {
            const int mcap = c->R * LL_FLAT_PER_RING / 10 + 16;
            const size_t vp_smem = ((size_t)mcap * mcap + (size_t)mcap * (VP_WORDS + 3)) * 4;
            if (!c->vp_attr_set) { LL_CUDA_CHECK(c, cudaFuncSetAttribute(k_odom_vote_partial, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vp_smem)); c->vp_attr_set = true; }
            LLProf pr(c, "k_odom_vote_partial");
            k_odom_vote_partial<<<dim3(10, n_lanes), VP_THREADS, vp_smem, s>>>(P, mcap);
        } else {
            LLProf pr(c, "k_odom_vote"); k_odom_vote<<<dim3(10, n_lanes), VOTE_THREADS, vote_smem, s>>>(P);
        }
        {
            LLProf pr(c, "k_lm_solve_odom");
            comm.seq_in = c->d_odom_seq[c->odom_comm_flip]; comm.seq_out = c->d_odom_seq[c->odom_comm_flip ^ (lm_parts > 1 ? 1 : 0)];
            if (lm_parts > 1) {   // the parts of a lane spin on each other's flags: cooperative launch = all resident or refused
                int a_n = c->B;
                void* args[] = {&P, &comm, &a_n};
                const void* fn = P.distortion ? (const void*)k_lm_solve_odom<true> : (const void*)k_lm_solve_odom<false>;
                LL_CUDA_CHECK(c, cudaLaunchCooperativeKernel(fn, dim3(n_lanes, lm_parts), dim3(lm_threads_used), args, 0, s));
                c->odom_comm_flip ^= 1;
            } else if (lm_cluster > 1) {
                if (P.distortion) LL_CUDA_CHECK(c, lm_launch_cluster(k_lm_solve_odom<true>, n_lanes, lm_cluster, lm_cluster_threads, s, P, comm, (int)c->B));
                else LL_CUDA_CHECK(c, lm_launch_cluster(k_lm_solve_odom<false>, n_lanes, lm_cluster, lm_cluster_threads, s, P, comm, (int)c->B));
            } else if (P.distortion) k_lm_solve_odom<true><<<n_lanes, lm_threads, 0, s>>>(P, comm, c->B);
            else k_lm_solve_odom<false><<<n_lanes, lm_threads, 0, s>>>(P, comm, c->B);
        }
        c->launches += 5;
    }
    if (P.distortion == 2) {
        LLProf pr(c, "k_odom_to_end");
        k_odom_to_end<<<dim3(32, n_lanes), 256, 0, s>>>(P, c->d_lsharp[0], c->d_lsharp[1], c->d_lflat[0], c->d_lflat[1]);
        c->launches += 1;
    }
    { LLProf pr(c, "k_odom_finalize"); k_odom_finalize<<<(n_lanes + 63) / 64, 64, 0, s>>>(c->d_lane, c->d_pose, c->d_status, n_lanes); }
    c->launches += 1;
    LL_CUDA_CHECK(c, cudaGetLastError());
    return LL_OK;
}

// used by the mapping module: grids over the gathered local map (slot 0 of the given buffers)
int ll_build_map_grid(ll_ctx* c, KnnGrid& g, const float4* pts, size_t lane_stride, int which, int n_lanes, int max_pts)
{
    return build_grid(c, g, pts, pts, lane_stride, which, n_lanes, max_pts);
}
