// Feature extraction on the device — replaces scanRegistration.cpp:100-377 (`laserCloudHandler` body).
//
//   k_classify       SR:109-110 (NaN / range filter), SR:139-169 (ring id), SR:177 (-atan2), per-tile ring histogram
//                    and stable ranks; coalesced, 1 thread / point
//   k_ring_scan      SR:215-221 (ring offsets = stable counting sort by ring), SR:178-193 (which point flips halfPassed)
//   k_scatter        SR:194-209 (relTime, intensity) + scatter into the ring-major cloud
//   k_ring_sort      SR:225-257: one CTA per (ring, lane): TMA bulk load of the ring into shared memory,
//                    11-tap curvature, six sector sorts in registers (one warp each) -> u16 sorted order per point
//   k_ring_pick      SR:251-359: greedy pick with +-5 suppression, one warp per ring (order-dependent part)
//   k_ring_lessflat  SR:361-376: less-flat collection and pcl::VoxelGrid(0.2) per ring
//   k_compact        concatenation of the per-ring outputs in the reference's publish order (SR:273-376)
//
// HBM-bound streaming work; no GEMM shape anywhere.  Algorithmic bytes per scan:
// 16 P (raw) + 16 P (full) read+write in the sort, 16 P read + 4 P + features written by k_ring_features.
#include <limits.h>
#include <math.h>
#include <stdlib.h>

#include "ll_ctx.h"
#include "ll_device.cuh"

namespace {

struct FeatParams {
    LaneState* lane;
    int8_t* ring8;
    uint8_t* rank8;
    int* tile_hist;
    float4* full;
    float* curv;
    int8_t* label;       // [B][Nmax] SR:234/272/278/324
    uint16_t* sorted16;  // [B][Nmax] per-sector sorted order + curvature class bits
    unsigned* brk;       // [B][R][brk_words] consecutive-gap break bits per ring
    int brk_words;
    float4* lf_tmp;
    int* ring_lists;
    int* ring_counts;
    float4* sharp;
    float4* flat;
    int* sharp_idx;
    int* lsharp_idx;
    int* flat_idx;
    float4* lsharp[2];
    float4* lflat[2];
    int Nmax, NT, R;
    int scan_line;
    float thres, lower_bound, up_bound, factor;
    float inv_leaf;   // 1.0f / 0.2f
    // ring sizes handled by this launch of the per-ring kernels: ring_lo < n <= 6 * SCAP + 11.  With the wide-sector
    // variant enabled the kernels run twice: <512> takes the rings of up to 3083 points, <1024> the longer ones (a CTA
    // whose ring belongs to the other variant leaves after two loads).  Rings longer than ring_cap are an error.
    int ring_lo, ring_cap;
    int* wide_list;   // rings longer than 3083 points as lane * R + ring, appended by k_ring_scan; [0] of wide_n = how many
    int* wide_n;
};

// SR:114-126: endOri from the last valid point and startOri
__device__ __forceinline__ float end_ori_of(float last_ori, float startOri)
{
    float endOri = (float)((double)last_ori + 2 * LL_PI);
    if ((double)(endOri - startOri) > 3 * LL_PI)
        endOri = (float)((double)endOri - 2 * LL_PI);
    else if ((double)(endOri - startOri) < LL_PI)
        endOri = (float)((double)endOri + 2 * LL_PI);
    return endOri;
}

__device__ __forceinline__ bool point_valid(float x, float y, float z, float thres)
{
    return isfinite(x) && isfinite(y) && isfinite(z) && !(x * x + y * y + z * z < thres * thres);  // SR:109 / SR:72
}
// SR:139 + SR:142-169: ring id of a valid point (-2 = outside the sensor's rings)
__device__ __forceinline__ int ring_of(const FeatParams& P, float x, float y, float z)
{
    // float atan / sqrt, * 180 in fp32, / M_PI in fp64, stored as float; then the ring formula.
    // Only the integer scanID survives, so it is first computed in plain fp32 (error << 1e-3 of a ring); the
    // literal sequence (atanf evaluated in fp64 and rounded once, double division) runs only when that value
    // lies within 2e-3 of a truncation boundary, where the roundings of the reference decide.
    const float t = z / sqrtf(x * x + y * y);
    const float angf = atanf(t) * 57.29578f;
    const float vf = P.scan_line == 16 ? (angf + 15.0f) * 0.5f + 0.5f : (P.scan_line == 32 ? (angf + 30.666666f) * 0.75f : (angf - P.lower_bound) * P.factor + 0.5f);
    const float fr = vf - floorf(vf);
    int scanID;
    if (vf > 1e-3f && fr > 2e-3f && fr < 1.0f - 2e-3f) {
        scanID = (int)vf;
        if (scanID >= P.scan_line) scanID = -2;
    } else {
        const float a = (float)atan((double)t);
        const float angle = (float)((double)(a * 180.0f) / LL_PI);
        if (P.scan_line == 16) {
            scanID = (int)((double)((angle + 15.0f) / 2.0f) + 0.5);
            if (scanID > 15 || scanID < 0) scanID = -2;
        } else if (P.scan_line == 32) {
            scanID = (int)(((double)angle + 92.0 / 3.0) * 3.0 / 4.0);
            if (scanID > 31 || scanID < 0) scanID = -2;
        } else {
            scanID = (int)((double)((angle - P.lower_bound) * P.factor) + 0.5);
            if (scanID >= 64 || scanID < 0) scanID = -2;
        }
    }
    return scanID;
}
// SR:177 ori = -atan2(y, x).  The reference's -atan2f is reproduced as the fp64 atan2 rounded once; only DECISIONS need that
// value to the last bit: the branch thresholds startOri - pi/2, startOri + 3pi/2, (ori - startOri) > pi (SR:180-192),
// endOri - 3pi/2 and endOri + pi/2 on ori + 2pi (SR:196-205), and the sign of relTime at ori = startOri (SR:207-208:
// int(intensity) is the ring id the odometry reads).  Modulo 2pi these are startOri + k pi/2 and endOri + pi/2, so a point
// farther than 1e-5 rad from all of them takes the fp32 atan2f (<= 2 ulp: its intensity moves by at most one ulp of the
// fraction, which nothing consumes).
__device__ __forceinline__ float scan_ori(float x, float y, float startOri, float endOri)
{
    float ori = -atan2f(y, x);
    const float qa = (ori - startOri) * 0.63661977f, qb = (ori - endOri) * 0.63661977f;   // in units of pi/2
    if (fabsf(qa - rintf(qa)) < 1e-5f || fabsf(qb - rintf(qb)) < 1e-5f) ori = -(float)atan2((double)y, (double)x);
    return ori;
}
// One CTA classifies CLS_TPB consecutive 256-point tiles of a lane: the points of all its tiles are requested before
// the first is used, so the dependent chain (lane state -> raw pointer -> point) is paid once per CTA, not per tile.
#define CLS_TPB 4
template <int MINB>
__global__ void __launch_bounds__(LL_TILE, MINB) k_classify(FeatParams P)
{
    __shared__ __align__(16) int cnt[2][LL_TILE / 32][LL_MAX_RINGS];   // per-warp ring counts of a tile, double-buffered over the tiles
    __shared__ int flip_s;
    reinterpret_cast<int4*>(&cnt[0][0][0])[threadIdx.x] = make_int4(0, 0, 0, 0);   // 2 x 8 x 64 ints = 256 int4
    if (threadIdx.x == 0) flip_s = INT_MAX;
    const int b = blockIdx.y;
    LaneState& L = P.lane[b];
    const int n = L.n_raw, sw = L.stride_words;
    const float startOri = L.start_ori, endOri = L.end_ori;   // found by k_reset_scan_state
    const float sdir_x = L.start_dir[0], sdir_y = L.start_dir[1];   // unit vector of the first valid point's azimuth
    const int half_seen = *(volatile int*)&L.half_idx;   // filter for the atomic below; a stale (larger) value only costs an atomic
    const int w = warp_id(), lane = lane_id();
    const int tile0 = blockIdx.x * CLS_TPB;
    float px[CLS_TPB], py[CLS_TPB], pz[CLS_TPB];
#pragma unroll
    for (int t = 0; t < CLS_TPB; ++t) {
        const int i = (tile0 + t) * LL_TILE + threadIdx.x;
        px[t] = py[t] = pz[t] = 0.f;
        if (i < n) {
            const uint32_t* p = L.raw + (size_t)i * sw;
            if (sw == 4) { const float4 v = __ldg(reinterpret_cast<const float4*>(p)); px[t] = v.x; py[t] = v.y; pz[t] = v.z; }  // slabs and the pool are 16-byte aligned
            else { px[t] = __uint_as_float(p[0]); py[t] = __uint_as_float(p[1]); pz[t] = __uint_as_float(p[2]); }
        }
    }
    __syncthreads();
    int fl_min = INT_MAX;
#pragma unroll
    for (int t = 0; t < CLS_TPB; ++t) {
        const int tile = tile0 + t;
        if (tile >= P.NT) break;
        int (*c)[LL_MAX_RINGS] = cnt[t & 1];
        const int i = tile * LL_TILE + threadIdx.x;
        const float x = px[t], y = py[t], z = pz[t];
        const bool valid = i < n && point_valid(x, y, z, P.thres);
        int ring = -1;
        bool flip = false;
        if (valid) {
            ring = ring_of(P, x, y, z);
            if (ring >= 0) {
                // Does this point flip halfPassed (SR:180-192 in the !halfPassed state: (ori - startOri) > pi after the
                // wrap into [startOri - pi/2, startOri + 3pi/2])?  With a = ori - startOri = (start azimuth) - (point azimuth)
                // that is a in (pi, 3pi/2]: sin a < 0 and cos a < 0, which two products with the start direction decide without
                // any arctangent.  Only points within 1e-3 rad of the quadrant's edges take the literal sequence (scan_ori).
                const float S = sdir_y * x - sdir_x * y, C = sdir_x * x + sdir_y * y, mm = 1e-6f * (x * x + y * y);
                if ((S > 0.f && S * S > mm) || (C > 0.f && C * C > mm)) flip = false;
                else if (S < 0.f && S * S > mm && C < 0.f && C * C > mm) flip = true;
                else {
                    float o = scan_ori(x, y, startOri, endOri);
                    if ((double)o < (double)startOri - LL_PI / 2)
                        o = (float)((double)o + 2 * LL_PI);
                    else if ((double)o > (double)startOri + LL_PI * 3 / 2)
                        o = (float)((double)o - 2 * LL_PI);
                    flip = (double)(o - startOri) > LL_PI;
                }
            }
        }
        P.ring8[(size_t)b * P.Nmax + i] = (int8_t)ring;
        if (flip) fl_min = min(fl_min, i);
        // stable rank inside the tile (SR:209 push_back order): warp match on the ring id, then prefix over the tile's warps
        const unsigned m = __match_any_sync(LL_FULL_MASK, ring);
        const int rank_in_warp = __popc(m & ((1u << lane) - 1u));
        if (ring >= 0 && rank_in_warp == 0) c[w][ring] = __popc(m);
        __syncthreads();
        if (threadIdx.x < P.R) {
            int run = 0;
            for (int ww = 0; ww < LL_TILE / 32; ++ww) { const int v = c[ww][threadIdx.x]; c[ww][threadIdx.x] = run; run += v; }
            P.tile_hist[((size_t)b * P.R + threadIdx.x) * P.NT + tile] = run;  // [lane][ring][tile]: scans run along tiles
        } else if (threadIdx.x >= 128) {
            // the other buffer (read last by the previous tile's rank write, before the barrier above) is cleared for the next tile
            reinterpret_cast<int4*>(&cnt[(t + 1) & 1][0][0])[threadIdx.x - 128] = make_int4(0, 0, 0, 0);   // 8 x 64 ints = 128 int4
        }
        __syncthreads();
        P.rank8[(size_t)b * P.Nmax + i] = ring >= 0 ? (uint8_t)(c[w][ring] + rank_in_warp) : 0;
    }
    fl_min = __reduce_min_sync(LL_FULL_MASK, fl_min);
    if (lane == 0 && fl_min != INT_MAX) atomicMin(&flip_s, fl_min);
    __syncthreads();
    // about half of all CTAs see a flip: one filtered atomic each
    if (threadIdx.x == 0 && flip_s < half_seen) atomicMin(&L.half_idx, flip_s);
}

// one CTA per lane, one warp per ring at a time: tile_hist[r][t] -> exclusive offset of tile t inside ring r;
// ring_begin[]; n_full.  Each lane owns four consecutive tiles per 128-tile chunk (one int4), all chunks of a ring
// are loaded before the first is scanned.
#define SCAN_MAX_CHUNKS 8   // 8 x 128 tiles x 256 points = 262144 points per scan on the vector path
__global__ void __launch_bounds__(1024) k_ring_scan(FeatParams P)
{
    __shared__ int tot[LL_MAX_RINGS];
    const int b = blockIdx.x;
    LaneState& L = P.lane[b];
    const int w = warp_id(), lane = lane_id();
    const int ntiles = (L.n_raw + LL_TILE - 1) / LL_TILE;
    const int nchunk = (ntiles + 127) / 128;
    for (int r = w; r < P.R; r += 32) {
        int* col = P.tile_hist + ((size_t)b * P.R + r) * P.NT;
        if (nchunk <= SCAN_MAX_CHUNKS && (P.NT & 3) == 0) {
            int4 v[SCAN_MAX_CHUNKS];
#pragma unroll
            for (int c = 0; c < SCAN_MAX_CHUNKS; ++c) {
                const int t = c * 128 + lane * 4;
                v[c] = make_int4(0, 0, 0, 0);
                if (c < nchunk && t < P.NT) v[c] = *reinterpret_cast<const int4*>(col + t);
                if (t + 0 >= ntiles) v[c].x = 0;
                if (t + 1 >= ntiles) v[c].y = 0;
                if (t + 2 >= ntiles) v[c].z = 0;
                if (t + 3 >= ntiles) v[c].w = 0;
            }
            int carry = 0;
#pragma unroll
            for (int c = 0; c < SCAN_MAX_CHUNKS; ++c) {
                if (c < nchunk) {
                    const int s = v[c].x + v[c].y + v[c].z + v[c].w;
                    int incl = s;
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(LL_FULL_MASK, incl, d); if (lane >= d) incl += o; }
                    int run = carry + incl - s;
                    int4 o;
                    o.x = run; run += v[c].x; o.y = run; run += v[c].y; o.z = run; run += v[c].z; o.w = run;
                    const int t = c * 128 + lane * 4;
                    if (t < P.NT) *reinterpret_cast<int4*>(col + t) = o;
                    carry += __shfl_sync(LL_FULL_MASK, incl, 31);
                }
            }
            if (lane == 0) tot[r] = carry;
        } else {
            const int per = (ntiles + 31) / 32;
            const int t0 = lane * per, t1 = min(t0 + per, ntiles);
            int s = 0;
            for (int t = t0; t < t1; ++t) s += col[t];
            int incl = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(LL_FULL_MASK, incl, d); if (lane >= d) incl += o; }
            int run = incl - s;
            for (int t = t0; t < t1; ++t) { const int c = col[t]; col[t] = run; run += c; }
            if (lane == 31) tot[r] = incl;
        }
    }
    __syncthreads();
    if (threadIdx.x < 32) {  // ring_begin = exclusive scan of the ring totals
        const int a0 = lane < P.R ? tot[lane] : 0, a1 = lane + 32 < P.R ? tot[lane + 32] : 0;
        int i0 = a0, i1 = a1;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int o0 = __shfl_up_sync(LL_FULL_MASK, i0, d), o1 = __shfl_up_sync(LL_FULL_MASK, i1, d);
            if (lane >= d) { i0 += o0; i1 += o1; }
        }
        const int t0 = __shfl_sync(LL_FULL_MASK, i0, 31), t1 = __shfl_sync(LL_FULL_MASK, i1, 31);
        if (lane < P.R) L.ring_begin[lane] = i0 - a0;
        if (lane + 32 < P.R) L.ring_begin[lane + 32] = t0 + i1 - a1;
        if (P.ring_cap > 6 * 512 + 11) {   // wide-sector variant enabled: hand it the rings the 512-key kernels leave alone
            if (lane < P.R && a0 > 6 * 512 + 11 && a0 <= P.ring_cap) P.wide_list[atomicAdd(P.wide_n, 1)] = b * P.R + lane;
            if (lane + 32 < P.R && a1 > 6 * 512 + 11 && a1 <= P.ring_cap) P.wide_list[atomicAdd(P.wide_n, 1)] = b * P.R + lane + 32;
        }
        if (lane == 0) {
            const int run = t0 + t1;
            L.ring_begin[P.R] = run;
            L.n_full = run;
            L.cur = L.last_slot ^ 1;  // this frame's less-sharp / less-flat go to the slot not holding the *Last clouds
            if (L.first_valid == INT_MAX) L.err = LL_E_EMPTY;
        }
    }
}

// CLS_TPB tiles per CTA as in k_classify: ring offsets in shared memory, all the CTA's loads issued up front
template <int MINB>
__global__ void __launch_bounds__(LL_TILE, MINB) k_scatter(FeatParams P)
{
    __shared__ int ring_begin_s[LL_MAX_RINGS];
    const int b = blockIdx.y;
    const LaneState& L = P.lane[b];
    const int n = L.n_raw, sw = L.stride_words, half_idx = L.half_idx;
    const float startOri = L.start_ori, endOri = L.end_ori;
    if (threadIdx.x < P.R) ring_begin_s[threadIdx.x] = L.ring_begin[threadIdx.x];
    const int tile0 = blockIdx.x * CLS_TPB;
    int ring[CLS_TPB], rank[CLS_TPB];
    float4 pt[CLS_TPB];
#pragma unroll
    for (int t = 0; t < CLS_TPB; ++t) {
        const int i = (tile0 + t) * LL_TILE + threadIdx.x;
        ring[t] = -1;
        if (i < n) {
            ring[t] = P.ring8[(size_t)b * P.Nmax + i];
            rank[t] = P.rank8[(size_t)b * P.Nmax + i];
            const uint32_t* p = L.raw + (size_t)i * sw;
            if (sw == 4) pt[t] = __ldg(reinterpret_cast<const float4*>(p));
            else pt[t] = make_float4(__uint_as_float(p[0]), __uint_as_float(p[1]), __uint_as_float(p[2]), 0.f);
        }
    }
    __syncthreads();
#pragma unroll
    for (int t = 0; t < CLS_TPB; ++t) {
        if (ring[t] < 0) continue;
        const int tile = tile0 + t, i = tile * LL_TILE + threadIdx.x;
        float ori = scan_ori(pt[t].x, pt[t].y, startOri, endOri);   // SR:177
        if (i <= half_idx) {  // SR:178-193 (the flipping point itself still takes this branch)
            if ((double)ori < (double)startOri - LL_PI / 2)
                ori = (float)((double)ori + 2 * LL_PI);
            else if ((double)ori > (double)startOri + LL_PI * 3 / 2)
                ori = (float)((double)ori - 2 * LL_PI);
        } else {  // SR:194-205
            ori = (float)((double)ori + 2 * LL_PI);
            if ((double)ori < (double)endOri - LL_PI * 3 / 2)
                ori = (float)((double)ori + 2 * LL_PI);
            else if ((double)ori > (double)endOri + LL_PI / 2)
                ori = (float)((double)ori - 2 * LL_PI);
        }
        const float relTime = (ori - startOri) / (endOri - startOri);             // SR:207
        const float intensity = (float)((double)ring[t] + 0.1 * (double)relTime);  // SR:208
        const int pos = ring_begin_s[ring[t]] + P.tile_hist[((size_t)b * P.R + ring[t]) * P.NT + tile] + rank[t];
        float4 o = pt[t];
        o.w = intensity;
        P.full[(size_t)b * P.Nmax + pos] = o;
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_ring_sort: one CTA per (ring, lane), one WARP per sector.  TMA bulk load of the ring slab into shared memory,
// 11-tap curvature (SR:225-235), then each warp sorts its sector (SR:257) entirely in registers: SCAP / 32 keys per
// lane, element e = r * 32 + lane, so compare-exchange distances >= 32 are register-to-register and the rest are
// shuffles — no shared-memory passes and no block barriers inside the sort.
// Keys are ONE 32-bit word: the high (32 - log2 SCAP) bits of the curvature's fp32 pattern (monotone for c >= 0) and
// the position inside the sector.  Curvatures that collide in the truncated bits end up adjacent (ordered by
// position); a fix-up pass re-sorts each such group (a handful per sector, two or three elements) on the full
// (curvature, position) key, which restores exactly the total order (curvature bits, index) of a 64-bit sort.
// The sorted order is written back as one u16 per point: local index | over(c > 0.1) << 14 | under(c < 0.1) << 15 —
// everything the greedy pick needs to know about the curvature.
// ---------------------------------------------------------------------------------------------------------
// Warp-wide ascending sort of NREG * 32 keys in registers, BLOCKED layout: element e = lane * NREG + q lives in register q
// of lane `lane`.  Compare-exchange distances below NREG stay inside a lane (plain min / max on registers); only the
// log2(32) largest distances of a merge cross lanes — 15 shuffle stages out of 45 for 512 keys (a striped layout needs
// 35), and shuffles issue at a quarter of the ALU rate.  The network is the all-ascending form of the bitonic sort:
// the first stage of every merge pairs e with its mirror image e ^ (kk - 1), the following ones e with e ^ j, and the
// lower index always keeps the minimum, so no stage needs a direction.
template <int NREG>
__device__ __forceinline__ void warp_sort_regs_blocked(unsigned (&k)[NREG], int lane)
{
    constexpr int N = NREG * 32;
#pragma unroll
    for (int kk = 2; kk <= N; kk <<= 1) {
        // mirror stage: partner = e ^ (kk - 1)
        if (kk <= NREG) {
#pragma unroll
            for (int q = 0; q < NREG; ++q) {
                const int pq = q ^ (kk - 1);
                if (q < pq) { const unsigned a = k[q], c = k[pq]; k[q] = min(a, c); k[pq] = max(a, c); }
            }
        } else {
            const int m = kk / NREG - 1;                       // lane distance mask: lane ^ m
            const bool lower = (lane & (kk / (2 * NREG))) == 0; // bit kk/2 of e
            unsigned o[NREG];
#pragma unroll
            for (int q = 0; q < NREG; ++q) o[q] = __shfl_xor_sync(LL_FULL_MASK, k[NREG - 1 - q], m);
#pragma unroll
            for (int q = 0; q < NREG; ++q) k[q] = lower ? min(k[q], o[q]) : max(k[q], o[q]);
        }
        // remaining stages of the merge: partner = e ^ j
#pragma unroll
        for (int j = kk >> 2; j > 0; j >>= 1) {
            if (j < NREG) {
#pragma unroll
                for (int q = 0; q < NREG; ++q) {
                    if ((q & j) == 0) { const unsigned a = k[q], c = k[q | j]; k[q] = min(a, c); k[q | j] = max(a, c); }
                }
            } else {
                const int m = j / NREG;
                const bool lower = (lane & m) == 0;
#pragma unroll
                for (int q = 0; q < NREG; ++q) {
                    const unsigned o = __shfl_xor_sync(LL_FULL_MASK, k[q], m);
                    k[q] = lower ? min(k[q], o) : max(k[q], o);
                }
            }
        }
    }
}

#define SORT_THREADS 192   // six sectors, one warp each
#define SORT_CHUNK 1536    // ring points staged per TMA copy (multiple of 32)
#define SORT_STAGE_BYTES ((SORT_CHUNK + 10) * 16)
template <int SCAP>
__device__ __forceinline__ void ring_sort_body(const FeatParams& P, const int r, const int b)
{
    constexpr int RCAP = 6 * SCAP + 16;
    constexpr int NTH = SORT_THREADS;
    constexpr int NREG = SCAP / 32;
    constexpr int IDXBITS = SCAP == 512 ? 9 : 10;
    constexpr unsigned IDXMASK = (1u << IDXBITS) - 1u;
    extern __shared__ __align__(128) unsigned char smem[];
    static_assert(6 * (SCAP + 32) * 4 <= SORT_STAGE_BYTES || SCAP > 512, "the sort keys reuse the staging buffer");
    constexpr size_t STAGE = (size_t)6 * (SCAP + 32) * 4 > SORT_STAGE_BYTES ? (size_t)6 * (SCAP + 32) * 4 : SORT_STAGE_BYTES;
    float4* pts = reinterpret_cast<float4*>(smem);                    // staging buffer of one chunk (+ halo) ...
    unsigned* keys = reinterpret_cast<unsigned*>(smem);               // ... reused for the sort keys afterwards
    float* curv_s = reinterpret_cast<float*>(smem + STAGE);
    int* sp = reinterpret_cast<int*>(curv_s + RCAP);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sp + 8);

    const int tid = threadIdx.x, lane = lane_id(), wid = warp_id();
    LaneState& L = P.lane[b];
    const int base = L.ring_begin[r], n = L.ring_begin[r + 1] - base, ntot = L.n_full;
    const float4* gfull = P.full + (size_t)b * P.Nmax;
    float* gcurv = P.curv + (size_t)b * P.Nmax;
    int8_t* glabel = P.label + (size_t)b * P.Nmax;
    uint16_t* gsorted = P.sorted16 + (size_t)b * P.Nmax;
    unsigned* gbrk = P.brk + ((size_t)b * P.R + r) * P.brk_words;

    auto global_curv = [&](int g) {  // the reference's stencil runs over ring joins (global index)
        float c = 0.f;
        if (g >= 5 && g < ntot - 5) {
            float dx = gfull[g - 5].x, dy = gfull[g - 5].y, dz = gfull[g - 5].z;
            for (int k = -4; k <= -1; ++k) { dx = dx + gfull[g + k].x; dy = dy + gfull[g + k].y; dz = dz + gfull[g + k].z; }
            dx = dx - 10 * gfull[g].x; dy = dy - 10 * gfull[g].y; dz = dz - 10 * gfull[g].z;
            for (int k = 1; k <= 5; ++k) { dx = dx + gfull[g + k].x; dy = dy + gfull[g + k].y; dz = dz + gfull[g + k].z; }
            c = dx * dx + dy * dy + dz * dz;
        }
        return c;
    };

    if (n - 11 < 6 || n > P.ring_cap) {  // SR:248 skip; or capacity overflow
        if (P.ring_lo == 0) {
            if (tid == 0 && n > P.ring_cap) L.err = LL_E_CAPACITY;
            for (int i = tid; i < n; i += NTH) { gcurv[base + i] = global_curv(base + i); glabel[base + i] = 0; }
        }
        return;
    }
    if (n <= P.ring_lo || n > 6 * SCAP + 11) return;   // the other variant's ring

    if (tid == 0) mbar_init(bar, 1);
    const int len = n - 11;
    if (tid <= 6) sp[tid] = 5 + len * tid / 6;  // SR:253-254 with ring-local indices
    __syncthreads();
    // The ring is staged through shared memory in chunks of SORT_CHUNK points (+ 5 points of halo on each side) so that
    // the CTA needs 37 KB instead of 75 KB: six CTAs per SM instead of three.  One TMA bulk copy per chunk; the
    // mbarrier's phase parity follows the chunk number.
    for (int c0 = 0, chunk = 0; c0 < n; c0 += SORT_CHUNK, ++chunk) {
        const int c1 = min(c0 + SORT_CHUNK, n);
        const int lo = max(c0 - 5, 0), hi = min(c1 + 5, n);
        if (tid == 0) tma_bulk_g2s(pts, gfull + base + lo, (uint32_t)(hi - lo) * 16u, bar);
        mbar_wait(bar, (uint32_t)(chunk & 1));
        const float4* pl = pts - lo;   // pl[i] = ring point i for lo <= i < hi
        for (int i = c0 + tid; i < c1; i += NTH) {
            float c;
            if (i >= 5 && i < n - 5) {
                float dx = pl[i - 5].x, dy = pl[i - 5].y, dz = pl[i - 5].z;
#pragma unroll
                for (int k = -4; k <= -1; ++k) { const float4 q = pl[i + k]; dx = dx + q.x; dy = dy + q.y; dz = dz + q.z; }
                { const float4 q = pl[i]; dx = dx - 10 * q.x; dy = dy - 10 * q.y; dz = dz - 10 * q.z; }
#pragma unroll
                for (int k = 1; k <= 5; ++k) { const float4 q = pl[i + k]; dx = dx + q.x; dy = dy + q.y; dz = dz + q.z; }
                c = dx * dx + dy * dy + dz * dz;
            } else {
                c = global_curv(base + i);
            }
            curv_s[i] = c;
            gcurv[base + i] = c;
            glabel[base + i] = 0;
        }
        // consecutive-gap break bits for the +-5 suppression (SR:290-293): bit i = |p[i] - p[i-1]|^2 > 0.05;
        // each warp covers 32 consecutive points (SORT_CHUNK is a multiple of 32), the ballot is the bitmask word
        const int w1 = c1 == n ? P.brk_words * 32 : c1;
        for (int i0 = c0 + (tid & ~31); i0 < w1; i0 += NTH) {
            const int i = i0 + lane;
            bool brk = false;
            if (i >= 1 && i < n) {
                const float4 a = pl[i], c = pl[i - 1];
                brk = (double)sqdist3(a.x, a.y, a.z, c.x, c.y, c.z) > 0.05;
            }
            const unsigned bits = __ballot_sync(LL_FULL_MASK, brk);
            if (lane == 0) gbrk[i0 >> 5] = bits;
        }
        __syncthreads();   // everybody is done with this chunk before the next copy (or the sort keys) overwrite it
    }

    // ---- one warp per sector (sectors tile [5, n-6)) ---------------------------------------------------------------
    const int j = wid, s0 = sp[j], slen = sp[j + 1] - s0;
    unsigned k[NREG];
#pragma unroll
    for (int q = 0; q < NREG; ++q) {
        // any assignment of the keys to the slots will do (the key carries its position): striped reads are conflict-free
        const int e = q * 32 + lane;
        k[q] = e < slen ? ((__float_as_uint(curv_s[s0 + e]) >> IDXBITS) << IDXBITS) | (unsigned)e : 0xFFFFFFFFu;
    }
    warp_sort_regs_blocked<NREG>(k, lane);
    // sorted rank rho = lane * NREG + q goes to sk[rho + rho / NREG]: one pad word per lane keeps the stores conflict-free
    unsigned* sk = keys + j * (SCAP + 32);
#define SK(e) sk[(e) + (e) / NREG]
#pragma unroll
    for (int q = 0; q < NREG; ++q) sk[lane * (NREG + 1) + q] = k[q];
    __syncwarp();
    // fix-up: groups of equal truncated curvature are re-sorted on (full curvature bits, position)
#pragma unroll 1
    for (int q = 0; q < NREG; ++q) {
        const int e = q * 32 + lane;
        if (e + 1 < slen) {
            const unsigned me = SK(e), nx = SK(e + 1);
            const bool head = (me >> IDXBITS) == (nx >> IDXBITS) && (e == 0 || (SK(e - 1) >> IDXBITS) != (me >> IDXBITS));
            if (head) {
                int g = 2;
                while (e + g < slen && (SK(e + g) >> IDXBITS) == (me >> IDXBITS)) ++g;
                for (int a = 1; a < g; ++a) {  // insertion sort of sk[e .. e+g)
                    const unsigned ka = SK(e + a);
                    const u64 fa = ((u64)__float_as_uint(curv_s[s0 + (int)(ka & IDXMASK)]) << 32) | (ka & IDXMASK);
                    int c = a - 1;
                    while (c >= 0) {
                        const unsigned kc = SK(e + c);
                        const u64 fc = ((u64)__float_as_uint(curv_s[s0 + (int)(kc & IDXMASK)]) << 32) | (kc & IDXMASK);
                        if (fc <= fa) break;
                        SK(e + c + 1) = kc;
                        --c;
                    }
                    SK(e + c + 1) = ka;
                }
            }
        }
    }
    __syncwarp();
    for (int e = lane; e < slen; e += 32) {
        const int idx = s0 + (int)(SK(e) & IDXMASK);
        const float c = curv_s[idx];
        gsorted[base + s0 + e] = (uint16_t)((unsigned)idx | ((double)c > 0.1 ? 0x4000u : 0u) | ((double)c < 0.1 ? 0x8000u : 0u));
    }
#undef SK
    __syncthreads();
    if (tid == 0) asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");   // the list form initialises it again for its next ring
}
// LIST = false: one CTA per (ring, lane).  LIST = true: a small fixed grid walks the list of long rings k_ring_scan made
// (normally empty: the launch then costs a few hundred CTAs that read one word and leave).
template <int SCAP, bool LIST>
__global__ void __launch_bounds__(SORT_THREADS) k_ring_sort(FeatParams P)
{
    if (!LIST) { ring_sort_body<SCAP>(P, blockIdx.x, blockIdx.y); return; }
    const int n = *P.wide_n;
    for (int e = blockIdx.x; e < n; e += gridDim.x) {
        const int v = P.wide_list[e];
        ring_sort_body<SCAP>(P, v % P.R, v / P.R);
        __syncthreads();
    }
}

// 64-bit window of a bit array held in 32-bit words
__device__ __forceinline__ u64 bits_at(const unsigned* words, int pos)
{
    const int w = pos >> 5;
    return (((u64)words[w + 1] << 32) | words[w]) >> (pos & 31);
}
// SR:288-311: +-5 neighbour suppression with the consecutive-gap break bits; one thread, bit operations only.
__device__ __forceinline__ void suppress_neighbours(const unsigned* brk, unsigned* picked, int ind)
{
    // forward: l = 1..5 stops at the first i = ind + l with brk[i]
    const unsigned f = (unsigned)bits_at(brk, ind + 1) & 0x1Fu;
    const int nf = f ? __ffs(f) - 1 : 5;
    // backward: l = -1..-5 tests gap(ind + l, ind + l + 1) = brk[ind + l + 1] -> brk[ind], brk[ind-1], ...
    const unsigned back = (unsigned)bits_at(brk, ind - 4) & 0x1Fu;   // bits: brk[ind-4 .. ind]
    const unsigned rev = __brev(back) >> 27;                          // bit k = brk[ind - k]
    const int nr = rev ? __ffs(rev) - 1 : 5;
    // the suppressed points form one contiguous range [ind - nr, ind + nf] (<= 11 bits, at most two words)
    const int lo = ind - nr, hi = ind + nf;
    const int w0 = lo >> 5, w1 = hi >> 5;
    const unsigned m_lo = 0xFFFFFFFFu << (lo & 31), m_hi = 0xFFFFFFFFu >> (31 - (hi & 31));
    if (w0 == w1) {
        picked[w0] |= m_lo & m_hi;
    } else {
        picked[w0] |= m_lo;
        picked[w1] |= m_hi;
    }
}

// the same as a range: the points a pick at `ind` suppresses are [lo, hi]
__device__ __forceinline__ void suppress_range(const unsigned* brk, int ind, int& lo, int& hi)
{
    const unsigned f = (unsigned)bits_at(brk, ind + 1) & 0x1Fu;
    const int nf = f ? __ffs(f) - 1 : 5;
    const unsigned back = (unsigned)bits_at(brk, ind - 4) & 0x1Fu;   // bits: brk[ind-4 .. ind]
    const unsigned rev = __brev(back) >> 27;                          // bit k = brk[ind - k]
    const int nr = rev ? __ffs(rev) - 1 : 5;
    lo = ind - nr; hi = ind + nf;
}
// sets bits [lo, hi] (at most 11 bits, two words) of a bitmask several lanes may update at once
__device__ __forceinline__ void mark_range(unsigned* picked, int lo, int hi)
{
    const int w0 = lo >> 5, w1 = hi >> 5;
    const unsigned m_lo = 0xFFFFFFFFu << (lo & 31), m_hi = 0xFFFFFFFFu >> (31 - (hi & 31));
    if (w0 == w1) atomicOr(&picked[w0], m_lo & m_hi);
    else { atomicOr(&picked[w0], m_lo); atomicOr(&picked[w1], m_hi); }
}

// ---------------------------------------------------------------------------------------------------------
// k_ring_pick: the greedy, order-dependent part of SR:251-359, one WARP per (ring, lane) so that thousands of
// rings run concurrently.  Per warp, shared memory holds the sorted candidates (u16), the gap-break bits and the
// picked bitmask, so the serial chain only touches shared memory; each ballot examines 32 sorted candidates.
// ---------------------------------------------------------------------------------------------------------
#define PICK_WARPS 4
template <int SCAP>
__device__ __forceinline__ void ring_pick_body(const FeatParams& P, const int r, const int b)
{
    constexpr int RCAP = 6 * SCAP + 16;
    constexpr int WORDS = (RCAP + 31) / 32 + 2;
    extern __shared__ __align__(16) unsigned char smem_pick[];
    const int lane = lane_id(), wid = warp_id();
    unsigned* picked = reinterpret_cast<unsigned*>(smem_pick) + (size_t)wid * (2 * WORDS + RCAP / 2);
    unsigned* brk = picked + WORDS;
    uint16_t* sorted = reinterpret_cast<uint16_t*>(brk + WORDS);
    LaneState& L = P.lane[b];
    const int base = L.ring_begin[r], n = L.ring_begin[r + 1] - base;
    int* my_counts = P.ring_counts + ((size_t)b * P.R + r) * 4;
    int* my_lists = P.ring_lists + ((size_t)b * P.R + r) * (LL_SHARP_PER_RING + LL_LSHARP_PER_RING + LL_FLAT_PER_RING);
    if (n - 11 < 6 || n > P.ring_cap) {
        if (lane == 0 && P.ring_lo == 0) { my_counts[0] = my_counts[1] = my_counts[2] = 0; }
        return;
    }
    if (n <= P.ring_lo || n > 6 * SCAP + 11) return;   // the other variant's ring
    const unsigned* gbrk = P.brk + ((size_t)b * P.R + r) * P.brk_words;
    const uint16_t* gsorted = P.sorted16 + (size_t)b * P.Nmax + base;
    for (int i = lane; i < WORDS; i += 32) { picked[i] = 0; brk[i] = i < P.brk_words ? gbrk[i] : 0u; }
    for (int i = lane; i < n; i += 32) sorted[i] = (i >= 5 && i < n - 6) ? gsorted[i] : (uint16_t)0;
    __syncwarp();
    int8_t* label = P.label + (size_t)b * P.Nmax + base;
    const int len = n - 11;
    int n_sharp = 0, n_lsharp = 0, n_flat = 0;
    for (int j = 0; j < 6; ++j) {
        const int s0 = 5 + len * j / 6, size = 5 + len * (j + 1) / 6 - s0;
        const uint16_t* seg = sorted + s0;
        // Both walks examine 32 sorted candidates per round and resolve the whole round in registers: every eligible
        // lane works out the range its pick would suppress, then the lanes are taken in sorted order - the first
        // eligible one is picked and every candidate inside its range drops out - which is exactly what the serial
        // loop does one candidate at a time.  Shared memory (the picked bits) is touched once per round.
        // corners: from the largest curvature down (SR:261-313)
        int npick = 0;
        for (int pos = size - 1; pos >= 0; pos -= 32) {
            const int p = pos - lane;
            const unsigned key = p >= 0 ? seg[p] : 0u;
            const int idx = (int)(key & 0x3FFFu);
            const bool over = p >= 0 && (key & 0x4000u);
            bool elig = over && !((picked[idx >> 5] >> (idx & 31)) & 1u);
            const unsigned nm = __ballot_sync(LL_FULL_MASK, p >= 0 && !over);
            int lo = 0, hi = -1;
            if (elig) suppress_range(brk, idx, lo, hi);
            int my_pick = 0;   // 1-based pick number of this lane's candidate
            bool full = false;
            unsigned em = __ballot_sync(LL_FULL_MASK, elig);
            while (em) {
                const int f = __ffs(em) - 1;
                if (npick + 1 > 20) { full = true; break; }  // SR:281-284: the 21st candidate ends the walk
                ++npick;
                if (lane == f) my_pick = npick;
                const int flo = __shfl_sync(LL_FULL_MASK, lo, f), fhi = __shfl_sync(LL_FULL_MASK, hi, f);
                if (idx >= flo && idx <= fhi) elig = false;   // includes the picked candidate itself
                em = __ballot_sync(LL_FULL_MASK, elig);
            }
            if (my_pick) {
                label[idx] = my_pick <= 2 ? 2 : 1;
                if (my_pick <= 2) my_lists[n_sharp + my_pick - 1] = base + idx;   // n_sharp / n_lsharp: counts before this sector
                my_lists[LL_SHARP_PER_RING + n_lsharp + my_pick - 1] = base + idx;
                mark_range(picked, lo, hi);
            }
            __syncwarp();
            if (full || nm) break;  // 20 picked, or reached curvature <= 0.1: nothing below can be picked
        }
        n_sharp += min(npick, 2);
        n_lsharp += npick;
        // flats: from the smallest curvature up (SR:316-359)
        int nsm = 0;
        for (int pos = 0; pos < size; pos += 32) {
            const int p = pos + lane;
            const unsigned key = p < size ? seg[p] : 0u;
            const int idx = (int)(key & 0x3FFFu);
            const bool under = p < size && (key & 0x8000u);
            bool elig = under && !((picked[idx >> 5] >> (idx & 31)) & 1u);
            const unsigned nm = __ballot_sync(LL_FULL_MASK, p < size && !under);
            int lo = 0, hi = -1;
            if (elig) suppress_range(brk, idx, lo, hi);
            int my_pick = 0;
            unsigned em = __ballot_sync(LL_FULL_MASK, elig);
            while (em && nsm < 4) {
                const int f = __ffs(em) - 1;
                ++nsm;
                if (lane == f) my_pick = nsm;
                const int flo = __shfl_sync(LL_FULL_MASK, lo, f), fhi = __shfl_sync(LL_FULL_MASK, hi, f);
                if (idx >= flo && idx <= fhi) elig = false;
                em = __ballot_sync(LL_FULL_MASK, elig);
            }
            if (my_pick) {
                label[idx] = -1;
                my_lists[LL_SHARP_PER_RING + LL_LSHARP_PER_RING + n_flat + my_pick - 1] = base + idx;
                if (my_pick < 4) mark_range(picked, lo, hi);   // SR:328-331: the 4th pick leaves before picked / suppression
            }
            __syncwarp();
            if (nsm >= 4 || nm) break;
        }
        n_flat += nsm;
        __syncwarp();
    }
    if (lane == 0) { my_counts[0] = n_sharp; my_counts[1] = n_lsharp; my_counts[2] = n_flat; }
}
template <int SCAP, bool LIST>
__global__ void __launch_bounds__(PICK_WARPS * 32) k_ring_pick(FeatParams P, int n_lanes)
{
    if (!LIST) {
        const int ring_lane = blockIdx.x * PICK_WARPS + warp_id();
        if (ring_lane < P.R * n_lanes) ring_pick_body<SCAP>(P, ring_lane % P.R, ring_lane / P.R);
        return;
    }
    const int n = *P.wide_n;
    for (int e = blockIdx.x * PICK_WARPS + warp_id(); e < n; e += gridDim.x * PICK_WARPS) {
        const int v = P.wide_list[e];
        ring_pick_body<SCAP>(P, v % P.R, v / P.R);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------------------
// k_ring_lessflat: less-flat = every sector point with label <= 0 (SR:361-367), then pcl::VoxelGrid(0.2) on
// the ring (SR:370-376): bbox -> voxel id -> groups of equal voxel id -> fp32 centroids accumulated in input order,
// emitted in ascending voxel id.  Along a ring consecutive points mostly fall into the same voxel, so the sort is
// done over RUNS of consecutive equal voxel ids (about a third of the points): keys (voxel id, run number), and the
// thread that owns a voxel walks its runs in order — the summation order stays the input order.
// Points and labels are read through L2; voxel ids, run starts and the sort keys live in shared memory.
// ---------------------------------------------------------------------------------------------------------
// the many-keys-per-thread variants live in their own functions so that their register needs do not spill the common path
template <int NS, int KPT>
__device__ __noinline__ void block_sort_big(u64* keys) { block_sort_u64_asc<NS, KPT>(keys); }
template <int NS, int KPT>
__device__ __noinline__ void block_sort_big32(unsigned* keys) { block_sort_asc<unsigned, NS, KPT>(keys); }
template <int SCAP>
__device__ __forceinline__ void ring_lessflat_body(const FeatParams& P, const int r, const int b)
{
    constexpr int RCAP = 6 * SCAP + 16;
    constexpr int KCAP = (RCAP <= 4096) ? 4096 : 8192;
    constexpr int NTH = 512;
    constexpr int CH = (RCAP + NTH - 1) / NTH;
    extern __shared__ __align__(128) unsigned char smem[];
    u64* keys = reinterpret_cast<u64*>(smem);                       // [KCAP] (voxel id << 32) | first compacted position of the run << 13 | its length
    int* vid = reinterpret_cast<int*>(keys + KCAP);                 // [RCAP] voxel id per less-flat point (compacted order)
    uint16_t* pidx = reinterpret_cast<uint16_t*>(vid + RCAP);       // [RCAP] ring-local index of that point
    int* ws = reinterpret_cast<int*>(pidx + RCAP + (RCAP & 1));
    float* red = reinterpret_cast<float*>(ws + 40);

    const int tid = threadIdx.x, lane = lane_id(), wid = warp_id();
    LaneState& L = P.lane[b];
    const int base = L.ring_begin[r], n = L.ring_begin[r + 1] - base;
    int* my_counts = P.ring_counts + ((size_t)b * P.R + r) * 4;
    if (n - 11 < 6 || n > P.ring_cap) { if (tid == 0 && P.ring_lo == 0) my_counts[3] = 0; return; }
    if (n <= P.ring_lo || n > 6 * SCAP + 11) return;   // the other variant's ring
    const float4* pts = P.full + (size_t)b * P.Nmax + base;
    const int8_t* label = P.label + (size_t)b * P.Nmax + base;

    // pass 1 (coalesced: round k, thread t owns point k * 512 + t): less-flat flag, floor(coordinate / leaf) as three
    // int16 parked in the (still unused) sort-key area, bounding box, per-(round, warp) counts for the compaction
    short4* cell = reinterpret_cast<short4*>(keys);   // [RCAP] x, y, z cell, w = 1: less-flat
    int* rcnt = reinterpret_cast<int*>(red + 112);    // [CH * 16 + 1] -> exclusive offsets
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    const float inv = P.inv_leaf;
    bool wide = false;   // a cell index that does not fit int16 (|coordinate| > 6.5 km at leaf 0.2)
#pragma unroll
    for (int k0 = 0; k0 < CH; k0 += 2) {
        // label and point of two rounds are requested together, whatever the label says: one round trip per pair of rounds
        int lb[2] = {1, 1};
        float4 qv[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = (k0 + u) * NTH + tid;
            qv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k0 + u < CH && i < n) { lb[u] = label[i]; qv[u] = pts[i]; }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int k = k0 + u, i = k * NTH + tid;
            if (k >= CH) break;
            bool lf = false;
            if (i < n) {
                lf = i >= 5 && i < n - 6 && lb[u] <= 0;
                short4 c4 = make_short4(0, 0, 0, 0);
                if (lf) {
                    const float4 q = qv[u];
                    mn[0] = fminf(mn[0], q.x); mn[1] = fminf(mn[1], q.y); mn[2] = fminf(mn[2], q.z);
                    mx[0] = fmaxf(mx[0], q.x); mx[1] = fmaxf(mx[1], q.y); mx[2] = fmaxf(mx[2], q.z);
                    const float f0 = floorf(q.x * inv), f1 = floorf(q.y * inv), f2 = floorf(q.z * inv);
                    wide |= !(fabsf(f0) < 32000.f && fabsf(f1) < 32000.f && fabsf(f2) < 32000.f);
                    c4 = make_short4((short)f0, (short)f1, (short)f2, 1);
                }
                cell[i] = c4;
            }
            const unsigned bl = __ballot_sync(LL_FULL_MASK, lf);
            if (lane == 0) rcnt[k * 16 + wid] = __popc(bl);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(LL_FULL_MASK, mn[a], d));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(LL_FULL_MASK, mx[a], d));
        }
        if (lane == 0) { red[a * 16 + wid] = mn[a]; red[(3 + a) * 16 + wid] = mx[a]; }
    }
    const int any_wide = __syncthreads_or(wide);
    if (tid < 6) {
        float v = red[tid * 16];
        for (int k = 1; k < NTH / 32; ++k) v = tid < 3 ? fminf(v, red[tid * 16 + k]) : fmaxf(v, red[tid * 16 + k]);
        red[96 + tid] = v;
    }
    if (wid == 1) {  // exclusive scan of the CH * 16 counts (round-major = input order)
        constexpr int NC = CH * 16, PER = (NC + 31) / 32;
        int v[PER], sum = 0;
#pragma unroll
        for (int q = 0; q < PER; ++q) { const int e = lane * PER + q; v[q] = e < NC ? rcnt[e] : 0; sum += v[q]; }
        int incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(LL_FULL_MASK, incl, d); if (lane >= d) incl += o; }
        int run = incl - sum;
#pragma unroll
        for (int q = 0; q < PER; ++q) { const int e = lane * PER + q; if (e < NC) rcnt[e] = run; run += v[q]; }
        if (lane == 31) rcnt[NC] = incl;
    }
    __syncthreads();
    const int m = rcnt[CH * 16];
    if (m == 0) { if (tid == 0) my_counts[3] = 0; return; }
    const float bmn[3] = {red[96], red[97], red[98]}, bmx[3] = {red[99], red[100], red[101]};
    const long long ddx = (long long)((bmx[0] - bmn[0]) * inv) + 1, ddy = (long long)((bmx[1] - bmn[1]) * inv) + 1,
                    ddz = (long long)((bmx[2] - bmn[2]) * inv) + 1;
    float4* gout = P.lf_tmp + (size_t)b * P.Nmax + base;
    const bool too_small = ddx * ddy * ddz > (long long)INT_MAX;   // PCL "leaf size too small": output = input
    const int min_b0 = (int)floorf(bmn[0] * inv), min_b1 = (int)floorf(bmn[1] * inv), min_b2 = (int)floorf(bmn[2] * inv);
    const int div0 = (int)floorf(bmx[0] * inv) - min_b0 + 1, div1 = (int)floorf(bmx[1] * inv) - min_b1 + 1;
    const int mul1 = div0, mul2 = div0 * div1;
    const long long vmax = (long long)mul2 * ((int)floorf(bmx[2] * inv) - min_b2 + 1);   // voxel ids are < vmax
    // pass 2: order-preserving compaction, voxel ids (from shared memory unless a cell index overflowed int16)
#pragma unroll
    for (int k = 0; k < CH; ++k) {
        const int i = k * NTH + tid;
        const short4 c4 = i < n ? cell[i] : make_short4(0, 0, 0, 0);
        const bool lf = c4.w != 0;
        const unsigned bl = __ballot_sync(LL_FULL_MASK, lf);
        if (lf) {
            const int o = rcnt[k * 16 + wid] + __popc(bl & ((1u << lane) - 1u));
            if (too_small) {
                gout[o] = pts[i];
            } else {
                int i0, i1, i2;
                if (!any_wide) {
                    i0 = (int)c4.x - min_b0; i1 = (int)c4.y - min_b1; i2 = (int)c4.z - min_b2;   // = (int)(floorf(x * inv) - (float)min_b), both exact integers
                } else {
                    const float4 q = pts[i];
                    i0 = (int)(floorf(q.x * inv) - (float)min_b0);
                    i1 = (int)(floorf(q.y * inv) - (float)min_b1);
                    i2 = (int)(floorf(q.z * inv) - (float)min_b2);
                }
                vid[o] = i0 + i1 * mul1 + i2 * mul2;
                pidx[o] = (uint16_t)i;
            }
        }
    }
    if (too_small) { if (tid == 0) my_counts[3] = m; return; }
    __syncthreads();
    // runs of consecutive equal voxel ids (threads own consecutive chunks of the compacted positions)
    const int MCH = (m + NTH - 1) / NTH;
    const int p0 = min(tid * MCH, m), p1 = min(p0 + MCH, m);
    int heads = 0;
    for (int p = p0; p < p1; ++p) heads += (p == 0 || vid[p] != vid[p - 1]);
    int R = 0;
    int ro = block_exclusive_scan(heads, ws, &R);
    int NS = 64;
    while (NS < R) NS <<= 1;
    // Keys of the run sort.  When the voxel id and the run number fit one 32-bit word together (most rings: the bounding
    // box of a ring on the ground is small) the sort runs on 32-bit keys = voxel id << log2(NS) | run number, and the
    // run's (start, length) waits in a side array; otherwise 64-bit keys carry (voxel id, start, length) themselves.
    const int bits_r = 31 - __clz(NS);
    const int bits_v = vmax <= 1 ? 1 : 64 - __clzll(vmax - 1);
    const bool fast = bits_v + bits_r <= 31 && NS <= KCAP;   // block-uniform; 31: the padding key 0xFFFFFFFF stays above every real key
    unsigned* key32 = reinterpret_cast<unsigned*>(keys);       // [NS] in the lower half of the key buffer
    unsigned* side = key32 + KCAP;                             // [R] in its upper half: run start << 13 | run length
    // every run is written by the thread that owns its head: the key when the run opens, (start, length) when the next head
    // (or, for the chunk's last run, the end of the run in a neighbour's chunk) closes it - no word is shared between threads
    {
        int open_ro = -1, open_p = 0;
        auto close_run = [&](int end) {
            const unsigned len = (unsigned)(end - open_p);
            if (fast) side[open_ro] = ((unsigned)open_p << 13) | len;
            else keys[open_ro] |= (u64)len;
        };
        for (int p = p0; p < p1; ++p)
            if (p == 0 || vid[p] != vid[p - 1]) {
                if (open_ro >= 0) close_run(p);
                if (fast) key32[ro] = ((unsigned)vid[p] << bits_r) | (unsigned)ro;
                else keys[ro] = ((u64)(unsigned)vid[p] << 32) | ((unsigned)p << 13);
                open_ro = ro; open_p = p;
                ++ro;
            }
        if (open_ro >= 0) {
            int e = p1;
            while (e < m && vid[e] == vid[open_p]) ++e;
            close_run(e);
        }
    }
    if (fast) { for (int i = R + tid; i < NS; i += NTH) key32[i] = 0xFFFFFFFFu; }
    else { for (int i = R + tid; i < NS; i += NTH) keys[i] = ~0ull; }
    __syncthreads();
    if (fast) {
        switch (NS) {   // all 512 threads call; NS / 8 (NS / 4) of them hold keys
            case 64: block_sort_asc<unsigned, 64, 2>(key32); break;
            case 128: block_sort_asc<unsigned, 128, 4>(key32); break;
            case 256: block_sort_asc<unsigned, 256, 8>(key32); break;
            case 512: block_sort_asc<unsigned, 512, 8>(key32); break;
            case 1024: block_sort_asc<unsigned, 1024, 8>(key32); break;
            case 2048: block_sort_asc<unsigned, 2048, 8>(key32); break;
            case 4096: block_sort_asc<unsigned, 4096, 8>(key32); break;
            default: block_sort_big32<8192, 16>(key32); break;
        }
    } else {
        switch (NS) {   // all 512 threads call; NS / 4 (or NS / 8) of them hold keys
            case 64: block_sort_u64_asc<64, 2>(keys); break;
            case 128: block_sort_u64_asc<128, 4>(keys); break;
            case 256: block_sort_u64_asc<256, 4>(keys); break;
            case 512: block_sort_u64_asc<512, 4>(keys); break;
            case 1024: block_sort_u64_asc<1024, 4>(keys); break;
            case 2048: block_sort_u64_asc<2048, 4>(keys); break;
            case 4096: block_sort_big<4096, 8>(keys); break;     // rings where nearly every point is its own run: rare
            default: block_sort_big<8192, 16>(keys); break;
        }
    }
    // voxels = groups of equal voxel id among the sorted runs; the group's first run's thread accumulates the centroid
    auto vox_of = [&](int q) { return fast ? key32[q] >> bits_r : (unsigned)(keys[q] >> 32); };
    const int RCH = (R + NTH - 1) / NTH;
    const int s0 = min(tid * RCH, R), s1 = min(s0 + RCH, R);
    int vheads = 0;
    for (int q = s0; q < s1; ++q) vheads += (q == 0 || vox_of(q) != vox_of(q - 1));
    int nvox = 0;
    int o = block_exclusive_scan(vheads, ws, &nvox);
    for (int q = s0; q < s1; ++q) {
        const unsigned v = vox_of(q);
        if (q == 0 || v != vox_of(q - 1)) {
            float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
            int cnt = 0;
            for (int g = q; g < R && vox_of(g) == v; ++g) {  // runs in ascending run number = input order
                const unsigned lo = fast ? side[key32[g] & (unsigned)(NS - 1)] : (unsigned)keys[g];
                const int e0 = (int)(lo >> 13), e1 = e0 + (int)(lo & 0x1FFFu);
#pragma unroll 4
                for (int p = e0; p < e1; ++p) {
                    const float4 a = pts[pidx[p]];
                    sx += a.x; sy += a.y; sz += a.z; si += a.w;
                }
                cnt += e1 - e0;
            }
            const float fc = (float)cnt;
            gout[o++] = make_float4(sx / fc, sy / fc, sz / fc, si / fc);
        }
    }
    if (tid == 0) my_counts[3] = nvox;
}
template <int SCAP, bool LIST>
__global__ void __launch_bounds__(512, 4) k_ring_lessflat(FeatParams P)
{
    if (!LIST) { ring_lessflat_body<SCAP>(P, blockIdx.x, blockIdx.y); return; }
    const int n = *P.wide_n;
    for (int e = blockIdx.x; e < n; e += gridDim.x) {
        const int v = P.wide_list[e];
        ring_lessflat_body<SCAP>(P, v % P.R, v / P.R);
        __syncthreads();
    }
}

// one CTA per (ring, lane): offsets = sums over the earlier rings; copy lists / clouds to their compact places
__global__ void __launch_bounds__(256) k_compact(FeatParams P)
{
    __shared__ int offs[4];
    const int r = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    LaneState& L = P.lane[b];
    const int* counts = P.ring_counts + (size_t)b * P.R * 4;
    if (tid < 32) {
        int s[4] = {0, 0, 0, 0};
        for (int q = tid; q < r; q += 32)
            for (int k = 0; k < 4; ++k) s[k] += counts[q * 4 + k];
        for (int k = 0; k < 4; ++k) {
            for (int d = 16; d > 0; d >>= 1) s[k] += __shfl_xor_sync(LL_FULL_MASK, s[k], d);
            if (tid == 0) offs[k] = s[k];
        }
    }
    __syncthreads();
    const int* lists = P.ring_lists + ((size_t)b * P.R + r) * (LL_SHARP_PER_RING + LL_LSHARP_PER_RING + LL_FLAT_PER_RING);
    const float4* gfull = P.full + (size_t)b * P.Nmax;
    const int cur = L.cur;
    const int ns = counts[r * 4 + 0], nls = counts[r * 4 + 1], nf = counts[r * 4 + 2], nlf = counts[r * 4 + 3];
    for (int k = tid; k < ns; k += blockDim.x) {
        const int g = lists[k];
        P.sharp[(size_t)b * P.R * LL_SHARP_PER_RING + offs[0] + k] = gfull[g];
        P.sharp_idx[(size_t)b * P.R * LL_SHARP_PER_RING + offs[0] + k] = g;
    }
    for (int k = tid; k < nls; k += blockDim.x) {
        const int g = lists[LL_SHARP_PER_RING + k];
        P.lsharp[cur][(size_t)b * P.R * LL_LSHARP_PER_RING + offs[1] + k] = gfull[g];
        P.lsharp_idx[(size_t)b * P.R * LL_LSHARP_PER_RING + offs[1] + k] = g;
    }
    for (int k = tid; k < nf; k += blockDim.x) {
        const int g = lists[LL_SHARP_PER_RING + LL_LSHARP_PER_RING + k];
        P.flat[(size_t)b * P.R * LL_FLAT_PER_RING + offs[2] + k] = gfull[g];
        P.flat_idx[(size_t)b * P.R * LL_FLAT_PER_RING + offs[2] + k] = g;
    }
    const float4* src = P.lf_tmp + (size_t)b * P.Nmax + L.ring_begin[r];
    float4* dst = P.lflat[cur] + (size_t)b * P.Nmax + offs[3];
    for (int k = tid; k < nlf; k += blockDim.x) dst[k] = src[k];
    if (tid == 0) {
        L.less_flat_ring_begin[r] = offs[3];
        if (r == P.R - 1) {
            L.n_sharp = offs[0] + ns;
            L.n_less_sharp = offs[1] + nls;
            L.n_flat = offs[2] + nf;
            L.n_less_flat = offs[3] + nlf;
            L.less_flat_ring_begin[P.R] = offs[3] + nlf;
        }
    }
}

// one warp per lane: per-scan state; startOri (SR:114) = azimuth of the first point that survives the filters (nearly
// always point 0) so that k_classify can decide where halfPassed flips (SR:178-193) in its single pass ...
__global__ void k_reset_scan_state(LaneState* lane, int n_lanes, float thres, int* wide_n)
{
    const int b = blockIdx.x, ln = lane_id();
    if (b >= n_lanes) return;
    if (b == 0 && ln == 0) *wide_n = 0;
    LaneState& L = lane[b];
    const int n = L.n_raw, sw = L.stride_words;
    int fv = -1;
    float sx = 0.f, sy = 0.f;
    for (int j0 = 0; j0 < n && fv < 0; j0 += 32) {
        const int j = j0 + ln;
        bool v = false;
        float x = 0.f, y = 0.f;
        if (j < n) {
            const uint32_t* p = L.raw + (size_t)j * sw;
            x = __uint_as_float(p[0]); y = __uint_as_float(p[1]);
            v = point_valid(x, y, __uint_as_float(p[2]), thres);
        }
        const unsigned m = __ballot_sync(LL_FULL_MASK, v);
        if (m) { fv = j0 + __ffs(m) - 1; sx = __shfl_sync(LL_FULL_MASK, x, __ffs(m) - 1); sy = __shfl_sync(LL_FULL_MASK, y, __ffs(m) - 1); }
    }
    // ... and endOri (SR:116-126) from the last one (nearly always point n - 1)
    int lv = -1;
    float ex = 0.f, ey = 0.f;
    for (int j0 = n - 1; j0 >= 0 && lv < 0 && fv >= 0; j0 -= 32) {
        const int j = j0 - ln;
        bool v = false;
        float x = 0.f, y = 0.f;
        if (j >= 0) {
            const uint32_t* p = L.raw + (size_t)j * sw;
            x = __uint_as_float(p[0]); y = __uint_as_float(p[1]);
            v = point_valid(x, y, __uint_as_float(p[2]), thres);
        }
        const unsigned m = __ballot_sync(LL_FULL_MASK, v);
        if (m) { lv = j0 - (__ffs(m) - 1); ex = __shfl_sync(LL_FULL_MASK, x, __ffs(m) - 1); ey = __shfl_sync(LL_FULL_MASK, y, __ffs(m) - 1); }
    }
    if (ln == 0) {
        L.first_valid = fv >= 0 ? fv : INT_MAX;
        L.last_valid = lv;
        const float so = fv >= 0 ? -(float)atan2((double)sy, (double)sx) : 0.f;   // = the ori k_classify stores for that point
        L.start_ori = so;
        { float sn, cs; sincosf(so, &sn, &cs); L.start_dir[0] = cs; L.start_dir[1] = -sn; }   // azimuth of the first valid point = -startOri
        L.end_ori = lv >= 0 ? end_ori_of(-(float)atan2((double)ey, (double)ex), so) : 0.f;
        L.half_idx = INT_MAX;
    }
}

}  // namespace

size_t ll_feature_smem_bytes(int SCAP)  // k_ring_sort: points + curvature + six sector key arrays + sector starts + mbarrier
{
    const int RCAP = 6 * SCAP + 16;
    const size_t keys = (size_t)6 * (SCAP + 32) * 4, stage = keys > (size_t)SORT_STAGE_BYTES ? keys : (size_t)SORT_STAGE_BYTES;
    return stage + (size_t)RCAP * 4 + 8 * 4 + 16;
}
size_t ll_lessflat_smem_bytes(int SCAP);
size_t ll_lessflat_smem_bytes(int SCAP)  // k_ring_lessflat: run sort keys, voxel ids, point indices, run starts, scratch
{
    const int RCAP = 6 * SCAP + 16, KCAP = RCAP <= 4096 ? 4096 : 8192;
    const int CH = (RCAP + 511) / 512;
    return (size_t)KCAP * 8 + (size_t)RCAP * 4 + (size_t)(RCAP + (RCAP & 1)) * 2 + 40 * 4 + 112 * 4 + (size_t)(CH * 16 + 4) * 4;
}

int ll_launch_features(ll_ctx* c, int n_lanes)
{
    FeatParams P;
    P.lane = c->d_lane; P.ring8 = c->d_ring8; P.rank8 = c->d_rank8; P.tile_hist = c->d_tile_hist;
    P.full = c->d_full; P.curv = c->d_curv; P.label = c->d_label; P.sorted16 = c->d_sorted16; P.brk = c->d_brk; P.brk_words = (c->RCAP + 31) / 32; P.lf_tmp = c->d_lf_tmp; P.ring_lists = c->d_ring_lists; P.ring_counts = c->d_ring_counts;
    P.sharp = c->d_sharp; P.flat = c->d_flat; P.sharp_idx = c->d_sharp_idx; P.lsharp_idx = c->d_lsharp_idx; P.flat_idx = c->d_flat_idx;
    P.lsharp[0] = c->d_lsharp[0]; P.lsharp[1] = c->d_lsharp[1]; P.lflat[0] = c->d_lflat[0]; P.lflat[1] = c->d_lflat[1];
    P.Nmax = c->Nmax; P.NT = c->NT; P.R = c->R; P.scan_line = c->cfg.scan_line;
    P.thres = c->cfg.minimum_range; P.lower_bound = c->cfg.lower_bound; P.up_bound = c->cfg.up_bound;
    P.factor = (c->cfg.scan_line - 1) / (c->cfg.up_bound - c->cfg.lower_bound);  // SR:441, fp32
    P.inv_leaf = 1.0f / 0.2f;
    P.wide_list = c->d_wide_list; P.wide_n = c->d_wide_n;
    P.ring_lo = 0; P.ring_cap = 6 * c->SCAP + 11;
    cudaStream_t s = c->stream;
    const dim3 tiles(c->NT, n_lanes);
    { LLProf pr(c, "k_reset_scan_state"); k_reset_scan_state<<<n_lanes, 32, 0, s>>>(c->d_lane, n_lanes, P.thres, c->d_wide_n); }
    const int cls_minb = getenv("LL_CLS_MINB") ? atoi(getenv("LL_CLS_MINB")) : 8, sct_minb = getenv("LL_SCT_MINB") ? atoi(getenv("LL_SCT_MINB")) : 8;
    {
        LLProf pr(c, "k_classify");
        const dim3 g((c->NT + CLS_TPB - 1) / CLS_TPB, n_lanes);
        if (cls_minb >= 8) k_classify<8><<<g, LL_TILE, 0, s>>>(P);
        else if (cls_minb >= 6) k_classify<6><<<g, LL_TILE, 0, s>>>(P);
        else if (cls_minb >= 5) k_classify<5><<<g, LL_TILE, 0, s>>>(P);
        else k_classify<4><<<g, LL_TILE, 0, s>>>(P);
    }
    { LLProf pr(c, "k_ring_scan"); k_ring_scan<<<n_lanes, 1024, 0, s>>>(P); }
    {
        LLProf pr(c, "k_scatter");
        const dim3 g((c->NT + CLS_TPB - 1) / CLS_TPB, n_lanes);
        if (sct_minb >= 8) k_scatter<8><<<g, LL_TILE, 0, s>>>(P);
        else if (sct_minb >= 6) k_scatter<6><<<g, LL_TILE, 0, s>>>(P);
        else if (sct_minb >= 5) k_scatter<5><<<g, LL_TILE, 0, s>>>(P);
        else k_scatter<4><<<g, LL_TILE, 0, s>>>(P);
    }
    const dim3 rings(c->R, n_lanes);
    const int pick_blocks = (c->R * n_lanes + PICK_WARPS - 1) / PICK_WARPS;
    auto smem_pick_of = [](int SCAP) { const int RCAP = 6 * SCAP + 16; return (size_t)PICK_WARPS * (2 * ((RCAP + 31) / 32 + 2) + RCAP / 2) * 4; };
    if (!c->feat_attr_set) {
        LL_CUDA_CHECK(c, cudaFuncSetAttribute(k_ring_sort<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ll_feature_smem_bytes(512)));
        LL_CUDA_CHECK(c, cudaFuncSetAttribute(k_ring_lessflat<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ll_lessflat_smem_bytes(512)));
        LL_CUDA_CHECK(c, cudaFuncSetAttribute(k_ring_pick<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pick_of(512)));
        LL_CUDA_CHECK(c, cudaFuncSetAttribute(k_ring_sort<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ll_feature_smem_bytes(1024)));
        LL_CUDA_CHECK(c, cudaFuncSetAttribute(k_ring_lessflat<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ll_lessflat_smem_bytes(1024)));
        LL_CUDA_CHECK(c, cudaFuncSetAttribute(k_ring_pick<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pick_of(1024)));
        c->feat_attr_set = true;
    }
    // rings of up to 3083 points: one warp sorts a sector of <= 512 keys in registers
    P.ring_lo = 0; P.ring_cap = 6 * c->SCAP + 11;
    { LLProf pr(c, "k_ring_sort"); k_ring_sort<512, false><<<rings, SORT_THREADS, ll_feature_smem_bytes(512), s>>>(P); }
    { LLProf pr(c, "k_ring_pick"); k_ring_pick<512, false><<<pick_blocks, PICK_WARPS * 32, smem_pick_of(512), s>>>(P, n_lanes); }
    { LLProf pr(c, "k_ring_lessflat"); k_ring_lessflat<512, false><<<rings, 512, ll_lessflat_smem_bytes(512), s>>>(P); }
    c->launches += 3;
    if (c->SCAP == 1024) {   // longer rings (up to 6155 points): the 1024-key variants; every other CTA leaves at once
        P.ring_lo = 6 * 512 + 11;
        const int wide_grid = 296;   // a fixed small grid over the (normally empty) list of long rings
        { LLProf pr(c, "k_ring_sort_wide"); k_ring_sort<1024, true><<<wide_grid, SORT_THREADS, ll_feature_smem_bytes(1024), s>>>(P); }
        { LLProf pr(c, "k_ring_pick_wide"); k_ring_pick<1024, true><<<wide_grid, PICK_WARPS * 32, smem_pick_of(1024), s>>>(P, n_lanes); }
        { LLProf pr(c, "k_ring_lessflat_wide"); k_ring_lessflat<1024, true><<<wide_grid, 512, ll_lessflat_smem_bytes(1024), s>>>(P); }
        c->launches += 3;
    }
    { LLProf pr(c, "k_compact"); k_compact<<<dim3(c->R, n_lanes), 256, 0, s>>>(P); }
    c->launches += 5;
    LL_CUDA_CHECK(c, cudaGetLastError());
    return LL_OK;
}
