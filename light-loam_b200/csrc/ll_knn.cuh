// Exact k-nearest-neighbour search over a hashed uniform grid — the replacement for
// pcl::KdTreeFLANN::nearestKSearch in the mapping module (LM:1882, LM:1948 with k = 5); the odometry's 1-NN
// (LO:494, LO:656) runs over the polar index of ll_odometry.cu and borrows the bucket streaming from here.
//
// The reference only ever accepts neighbours under a fixed radius (d2[4] < 1.0, LM:1884/1952), so a radius-bounded grid search returns exactly what the kd-tree would, provided the fp32
// distance arithmetic (FLANN L2_Simple: ((dx*dx)+(dy*dy))+(dz*dz)) and the tie rule (lowest target index)
// are pinned.  Points are bucketed by hash(cell) without storing cell keys: colliding cells only add
// candidates that are distance-tested anyway, so the result stays exact.
//
// One warp per query.  Shell s = all cells at Chebyshev distance s from the query's cell; every point closer
// than s*h is inside shells 0..s, so the search stops as soon as the k-th best is closer than s*h - eps.
// Within a shell the 32 lanes first fetch 32 bucket headers at once, then stream the concatenation of the
// bucket ranges with coalesced float4 loads.
#pragma once
#include "ll_device.cuh"

struct GridView {
    const int* start;      // [T+1] for this lane
    const float4* sorted;  // bucket-ordered points, .w bits = original index (low 24) | ring id (high 8)
    int Tmask;
    float h, inv_h;
};

template <int K>
struct WarpKnn {
    u64 key[K];  // per-lane sorted ascending: (d2 bits << 32) | original index
    __device__ __forceinline__ void init()
    {
#pragma unroll
        for (int i = 0; i < K; ++i) key[i] = ~0ull;
    }
    __device__ __forceinline__ void insert(u64 k)
    {
        if (k >= key[K - 1]) return;
        if (K > 1) {  // a bucket can be visited twice (hash collisions across shells): keep entries distinct
#pragma unroll
            for (int i = 0; i < K - 1; ++i) if (key[i] == k) return;
        }
        key[K - 1] = k;
#pragma unroll
        for (int i = K - 1; i > 0; --i) {
            if (key[i] < key[i - 1]) { const u64 t = key[i]; key[i] = key[i - 1]; key[i - 1] = t; }
        }
    }
};

// Merges the lanes' lists into the K smallest distinct keys (warp-uniform result in out[]).
template <int K>
__device__ __forceinline__ void warp_merge_topk(const WarpKnn<K>& loc, u64 out[K])
{
    int head = 0;
#pragma unroll
    for (int r = 0; r < K; ++r) {
        u64 mine = ~0ull;
#pragma unroll
        for (int i = 0; i < K; ++i) if (i == head) mine = loc.key[i];
        const u64 m = warp_min_u64(mine);
        out[r] = m;
        if (mine == m && m != ~0ull) ++head;  // every lane holding this (d2, idx) drops it: duplicates collapse
    }
}

// Streams the points of up to 32 buckets (one per lane, negative = none; duplicates must already be removed) through
// f(float4 point): bucket ranges are concatenated across the warp and read with coalesced float4 loads, two chunks
// of 32 candidates in flight per iteration.
template <typename F>
__device__ __forceinline__ void grid_stream_ranges(const float4* __restrict__ sorted, int beg, int cnt, F&& f);
template <typename F>
__device__ __forceinline__ void grid_stream_buckets(const GridView& g, int bucket, F&& f)
{
    int beg = 0, cnt = 0;
    if (bucket >= 0) {
        beg = g.start[bucket];
        cnt = g.start[bucket + 1] - beg;
    }
    grid_stream_ranges(g.sorted, beg, cnt, f);
}
// The same for explicit ranges: lane l contributes sorted[beg, beg + cnt) (cnt may be 0).
template <typename F>
__device__ __forceinline__ void grid_stream_ranges(const float4* __restrict__ sorted, int beg, int cnt, F&& f)
{
    const int lane = lane_id();
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(LL_FULL_MASK, incl, d); if (lane >= d) incl += o; }
    const int excl = incl - cnt;
    const int total = __shfl_sync(LL_FULL_MASK, incl, 31);
    for (int t0 = 0; t0 < total; t0 += 64) {
        float4 pv[2];
        bool ok[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int t = t0 + u * 32 + lane;
            ok[u] = false;
            if (t0 + u * 32 < total) {  // warp-uniform
                int lo = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1) {
                    const int cand = lo + step;
                    const int pv_ = __shfl_sync(LL_FULL_MASK, excl, cand & 31);
                    if (cand < 32 && pv_ <= t) lo = cand;
                }
                const int cbeg = __shfl_sync(LL_FULL_MASK, beg, lo);
                const int cexc = __shfl_sync(LL_FULL_MASK, excl, lo);
                ok[u] = t < total;
                if (ok[u]) pv[u] = sorted[cbeg + (t - cexc)];
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u)
            if (ok[u]) f(pv[u]);
    }
}

// Streams every point of shell s (cells at Chebyshev distance s from (cx,cy,cz); s == 1 also covers s == 0)
// through f(float4 point), 32 bucket headers per round.
template <typename F>
__device__ __forceinline__ void grid_visit_shell(const GridView& g, int cx, int cy, int cz, int s, F&& f)
{
    const int lane = lane_id();
    const int w = 2 * s + 1, ncell = w * w * w;
    for (int e0 = 0; e0 < ncell; e0 += 32) {
        const int e = e0 + lane;
        int bucket = -1 - lane;  // distinct invalid ids so match_any never groups invalid lanes with valid ones
        if (e < ncell) {
            const int dx = e % w - s, dy = (e / w) % w - s, dz = e / (w * w) - s;
            const int cheb = max(abs(dx), max(abs(dy), abs(dz)));
            if (cheb == s || s == 1) bucket = cell_bucket(cx + dx, cy + dy, cz + dz, g.Tmask);
        }
        const unsigned grp = __match_any_sync(LL_FULL_MASK, bucket);
        if ((__ffs(grp) - 1) != lane) bucket = -1;
        grid_stream_buckets(g, bucket, f);
    }
}

// Returns (warp-uniform) the K best keys; the caller applies its own d2 cutoff.
// r_search: every neighbour with true distance < r_search is guaranteed to be considered.
template <int K>
__device__ __forceinline__ void grid_knn(const GridView& g, float qx, float qy, float qz, float r_search, u64 out[K])
{
    const float eps = 1e-3f;
    const int cx = (int)floorf(qx * g.inv_h), cy = (int)floorf(qy * g.inv_h), cz = (int)floorf(qz * g.inv_h);
    WarpKnn<K> loc;
    loc.init();
    const int smax = (int)ceilf((r_search + eps) * g.inv_h);
    for (int s = 1; s <= smax; ++s) {  // the first pass covers shells 0 and 1 (27 cells) at once
        grid_visit_shell(g, cx, cy, cz, s, [&](const float4 p) {
            const float d2 = sqdist3(qx, qy, qz, p.x, p.y, p.z);
            loc.insert(((u64)__float_as_uint(d2) << 32) | ((unsigned)__float_as_int(p.w) & 0xFFFFFFu));
        });
        warp_merge_topk<K>(loc, out);
        if (out[K - 1] != ~0ull) {
            const float kth = __uint_as_float((unsigned)(out[K - 1] >> 32));
            const float safe = (float)s * g.h - eps;
            if (kth < safe * safe) break;
        }
    }
}
