// Stable LSD radix sort of (64-bit key, 32-bit value) pairs and an exclusive prefix sum over int arrays — the two
// device-wide primitives of the map's voxel filter (pcl::VoxelGrid restated: LM:1814-1822, LM:2155-2168), hand-written
// so that nothing on the scan-to-map path is a library call.
//
// One pass per 8-bit digit, three kernels per pass:
//   k_rs_count    per tile of 2048 elements: how many keys carry each digit value -> hist[digit][tile]
//   k_rs_offsets  one CTA per digit value: exclusive prefix of its row over the tiles, and the row total
//   k_rs_scatter  digit bases from the 256 totals; the tile recomputes its digits and moves every pair to
//                 base[digit] + row prefix[tile] + stable rank inside the tile
// Stability inside a tile: warp w owns 256 consecutive elements and walks them 32 at a time; `match_any` ranks the lanes of
// a chunk that share a digit, a per-warp running counter carries the rank from chunk to chunk, and a prefix over the
// warps' counters (digit-major) orders the warps.  Only the digits that can differ are sorted (the caller lists them).
//
// The prefix sum is the usual three-level scheme: block sums of 2048-int chunks, a recursive scan of the block sums, and
// a second sweep that adds each chunk's base.
#pragma once
#include <cuda_runtime.h>

#include "ll_device.cuh"

#define RS_THREADS 256
#define RS_ITEMS 8
#define RS_TILE (RS_THREADS * RS_ITEMS)   // 2048 elements per CTA
#define RS_BINS 256

namespace llsort {

// Sizes live on the device: the arrays are allocated (and the grids sized) for the capacity, the number of elements in use
// is read from *n_dev by every kernel, and the CTAs beyond it leave at once - a small map costs a small sort.
// LenSpec: the length a scan kernel works on = *n_dev divided `level` times by the scan chunk (the chunk-sum levels).
struct LenSpec { const int* n_dev; int level; };
__device__ __forceinline__ int rs_tiles(int n) { return (n + RS_TILE - 1) / RS_TILE; }
__device__ __forceinline__ long long len_of(const LenSpec L)
{
    long long n = *L.n_dev;
    for (int k = 0; k < L.level; ++k) n = (n + 2047) / 2048;
    return n;
}

// per-warp digit counts of one tile; keys[] = the thread's 8 keys (chunk c of warp w = elements w * 256 + c * 32 + lane)
__device__ __forceinline__ void rs_load_and_count(const u64* __restrict__ keys_in, long long n, int shift, long long tile0, u64 (&k)[RS_ITEMS],
                                                  int (*warp_cnt)[RS_BINS])
{
    const int w = warp_id(), lane = lane_id();
    const long long base = tile0 + (long long)w * (32 * RS_ITEMS);
#pragma unroll
    for (int c = 0; c < RS_ITEMS; ++c) {
        const long long i = base + c * 32 + lane;
        k[c] = i < n ? keys_in[i] : ~0ull;
    }
#pragma unroll
    for (int c = 0; c < RS_ITEMS; ++c) {
        const long long i = base + c * 32 + lane;
        const int d = i < n ? (int)((k[c] >> shift) & (RS_BINS - 1)) : -1 - lane;   // distinct invalid ids: never grouped
        const unsigned m = __match_any_sync(LL_FULL_MASK, d);
        if (d >= 0 && lane == __ffs(m) - 1) warp_cnt[w][d] += __popc(m);   // one lane per distinct digit of the chunk; the warp owns its row
        __syncwarp();
    }
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_count(const u64* __restrict__ keys_in, const int* __restrict__ n_dev, int shift, int* __restrict__ hist)
{
    __shared__ int warp_cnt[RS_THREADS / 32][RS_BINS];
    const int n = *n_dev, ntiles = rs_tiles(n);
    if ((int)blockIdx.x >= ntiles) return;
    for (int q = threadIdx.x; q < (RS_THREADS / 32) * RS_BINS; q += RS_THREADS) (&warp_cnt[0][0])[q] = 0;
    __syncthreads();
    u64 k[RS_ITEMS];
    rs_load_and_count(keys_in, n, shift, (long long)blockIdx.x * RS_TILE, k, warp_cnt);
    __syncthreads();
    int s = 0;
#pragma unroll
    for (int w = 0; w < RS_THREADS / 32; ++w) s += warp_cnt[w][threadIdx.x];   // RS_THREADS == RS_BINS: thread d sums digit d
    hist[(size_t)threadIdx.x * ntiles + blockIdx.x] = s;
}

// hist[d][0..ntiles) -> its exclusive prefix, totals[d] = the row sum; grid = 256 (one CTA per digit value)
__global__ void __launch_bounds__(RS_THREADS) k_rs_offsets(int* __restrict__ hist, const int* __restrict__ n_dev, int* __restrict__ totals)
{
    __shared__ int ws[40];
    const int ntiles = rs_tiles(*n_dev), d = blockIdx.x;
    int* row = hist + (size_t)d * ntiles;
    int running = 0;
    for (int base = 0; base < ntiles; base += RS_THREADS * 4) {
        const int i0 = base + threadIdx.x * 4;
        int v[4], sum = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) { v[q] = i0 + q < ntiles ? row[i0 + q] : 0; sum += v[q]; }
        int tot = 0;
        int run = running + block_exclusive_scan(sum, ws, &tot);
#pragma unroll
        for (int q = 0; q < 4; ++q) { if (i0 + q < ntiles) row[i0 + q] = run; run += v[q]; }
        running += tot;
    }
    if (threadIdx.x == 0) totals[d] = running;
}

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const u64* __restrict__ keys_in, const int* __restrict__ vals_in, u64* __restrict__ keys_out,
                                                          int* __restrict__ vals_out, const int* __restrict__ n_dev, int shift, const int* __restrict__ hist_scanned,
                                                          const int* __restrict__ totals)
{
    __shared__ int warp_cnt[RS_THREADS / 32][RS_BINS];
    __shared__ int ws[40];
    const int n = *n_dev, ntiles = rs_tiles(n);
    if ((int)blockIdx.x >= ntiles) return;
    const int dbase = block_exclusive_scan(totals[threadIdx.x], ws, nullptr);   // keys with a smaller digit, all tiles (thread d = digit d)
    for (int q = threadIdx.x; q < (RS_THREADS / 32) * RS_BINS; q += RS_THREADS) (&warp_cnt[0][0])[q] = 0;
    __syncthreads();
    const long long tile0 = (long long)blockIdx.x * RS_TILE;
    u64 k[RS_ITEMS];
    rs_load_and_count(keys_in, n, shift, tile0, k, warp_cnt);
    __syncthreads();
    {   // warp_cnt[w][d] := where warp w's first key with digit d goes = tile's start for d + counts of the warps before w
        const int d = threadIdx.x;
        int run = dbase + hist_scanned[(size_t)d * ntiles + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_THREADS / 32; ++w) { const int c = warp_cnt[w][d]; warp_cnt[w][d] = run; run += c; }
    }
    __syncthreads();
    const int w = warp_id(), lane = lane_id();
    const long long base = tile0 + (long long)w * (32 * RS_ITEMS);
#pragma unroll
    for (int c = 0; c < RS_ITEMS; ++c) {
        const long long i = base + c * 32 + lane;
        const int d = i < n ? (int)((k[c] >> shift) & (RS_BINS - 1)) : -1 - lane;
        const unsigned m = __match_any_sync(LL_FULL_MASK, d);
        if (d >= 0) {
            const int pos = warp_cnt[w][d] + __popc(m & ((1u << lane) - 1u));
            keys_out[pos] = k[c];
            vals_out[pos] = vals_in[i];
        }
        __syncwarp();
        if (d >= 0 && lane == __ffs(m) - 1) warp_cnt[w][d] += __popc(m);
        __syncwarp();
    }
}

// ---- exclusive prefix sum over ints -------------------------------------------------------------------------------------
#define SC_CHUNK 2048   // ints per CTA (256 threads x 8)
__global__ void __launch_bounds__(256) k_scan_chunks(const int* __restrict__ in, int* __restrict__ out, int* __restrict__ sums, LenSpec L)
{
    __shared__ int ws[40];
    const long long n = len_of(L);
    if ((long long)blockIdx.x * SC_CHUNK >= n) return;   // chunks past the data: their sums are never read by a live chunk
    const long long base = (long long)blockIdx.x * SC_CHUNK + (long long)threadIdx.x * 8;
    int v[8], s = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) { v[q] = base + q < n ? in[base + q] : 0; s += v[q]; }
    int tot = 0;
    int run = block_exclusive_scan(s, ws, &tot);
#pragma unroll
    for (int q = 0; q < 8; ++q) { if (base + q < n) out[base + q] = run; run += v[q]; }
    if (threadIdx.x == 0 && sums) sums[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(256) k_scan_add(int* __restrict__ out, const int* __restrict__ chunk_base, LenSpec L)
{
    const long long n = len_of(L);
    if ((long long)blockIdx.x * SC_CHUNK >= n) return;
    const int add = chunk_base[blockIdx.x];
    const long long base = (long long)blockIdx.x * SC_CHUNK + (long long)threadIdx.x * 8;
#pragma unroll
    for (int q = 0; q < 8; ++q) if (base + q < n) out[base + q] += add;
}

// scratch ints needed by scan_exclusive for a capacity of n elements (all levels of chunk sums)
static inline size_t scan_scratch_ints(long long n)
{
    size_t total = 0;
    while (n > SC_CHUNK) { n = (n + SC_CHUNK - 1) / SC_CHUNK; total += (size_t)n + 8; }
    return total + 8;
}
// out[i] = sum of in[0..i) for i < len_of(L) (in == out allowed); n_cap = the capacity the launches are sized for.
// Returns the kernels launched.
static inline int scan_exclusive(const int* in, int* out, long long n_cap, LenSpec L, int* scratch, cudaStream_t s)
{
    if (n_cap <= 0) return 0;
    const long long nchunks = (n_cap + SC_CHUNK - 1) / SC_CHUNK;
    if (nchunks == 1) { k_scan_chunks<<<1, 256, 0, s>>>(in, out, nullptr, L); return 1; }
    k_scan_chunks<<<(unsigned)nchunks, 256, 0, s>>>(in, out, scratch, L);
    LenSpec up = L; up.level += 1;
    int launches = 1 + scan_exclusive(scratch, scratch, nchunks, up, scratch + nchunks + 8, s);
    k_scan_add<<<(unsigned)nchunks, 256, 0, s>>>(out, scratch, L);
    return launches + 1;
}

// hist ints needed for a capacity of n elements: the (digit, tile) table + the 256 row totals
static inline size_t sort_hist_ints(long long n) { return (size_t)((n + RS_TILE - 1) / RS_TILE) * RS_BINS + RS_BINS; }

// Sorts the first *n_dev pairs by the key bits covered by `shifts` (8-bit digits, least significant first), stable.
// keys[0] / vals[0] hold the input; returns the index (0 or 1) of the buffer pair holding the result.  n_cap = capacity
// (sizes the grids); hist = sort_hist_ints(n_cap) ints.  *launches += kernels launched.
static inline int sort_pairs(u64* keys[2], int* vals[2], long long n_cap, const int* n_dev, const int* shifts, int n_shifts, int* hist, cudaStream_t s, int* launches)
{
    const int ntiles = (int)((n_cap + RS_TILE - 1) / RS_TILE);
    int* totals = hist + (size_t)ntiles * RS_BINS;
    int cur = 0;
    for (int p = 0; p < n_shifts; ++p) {
        k_rs_count<<<ntiles, RS_THREADS, 0, s>>>(keys[cur], n_dev, shifts[p], hist);
        k_rs_offsets<<<RS_BINS, RS_THREADS, 0, s>>>(hist, n_dev, totals);
        k_rs_scatter<<<ntiles, RS_THREADS, 0, s>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n_dev, shifts[p], hist, totals);
        if (launches) *launches += 3;
        cur ^= 1;
    }
    return cur;
}

}  // namespace llsort
