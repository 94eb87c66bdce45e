// Stable LSD radix sort of (64-bit key, 32-bit value) pairs and an exclusive prefix sum over int arrays — the two
// device-wide primitives of the map's voxel filter (pcl::VoxelGrid restated: LM:1814-1822, LM:2155-2168), hand-written
// so that nothing on the scan-to-map path is a library call.
//
// One pass per 8-bit digit, three kernels per pass:
//   k_rs_count    per tile of 2048 elements: how many keys carry each digit value -> hist[digit][tile]
//   (prefix sum)  exclusive scan of hist in (digit, tile) order = where each tile's run of each digit value starts
//   k_rs_scatter  the tile recomputes its digits and moves every pair to  start + stable rank inside the tile
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

__global__ void __launch_bounds__(RS_THREADS) k_rs_count(const u64* __restrict__ keys_in, long long n, int shift, int* __restrict__ hist, int ntiles)
{
    __shared__ int warp_cnt[RS_THREADS / 32][RS_BINS];
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

__global__ void __launch_bounds__(RS_THREADS) k_rs_scatter(const u64* __restrict__ keys_in, const int* __restrict__ vals_in, u64* __restrict__ keys_out,
                                                          int* __restrict__ vals_out, long long n, int shift, const int* __restrict__ hist_scanned, int ntiles)
{
    __shared__ int warp_cnt[RS_THREADS / 32][RS_BINS];
    for (int q = threadIdx.x; q < (RS_THREADS / 32) * RS_BINS; q += RS_THREADS) (&warp_cnt[0][0])[q] = 0;
    __syncthreads();
    const long long tile0 = (long long)blockIdx.x * RS_TILE;
    u64 k[RS_ITEMS];
    rs_load_and_count(keys_in, n, shift, tile0, k, warp_cnt);
    __syncthreads();
    {   // warp_cnt[w][d] := where warp w's first key with digit d goes = tile's start for d + counts of the warps before w
        const int d = threadIdx.x;
        int run = hist_scanned[(size_t)d * ntiles + blockIdx.x];
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
__global__ void __launch_bounds__(256) k_scan_chunks(const int* __restrict__ in, int* __restrict__ out, int* __restrict__ sums, long long n)
{
    __shared__ int ws[40];
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
__global__ void __launch_bounds__(256) k_scan_add(int* __restrict__ out, const int* __restrict__ chunk_base, long long n)
{
    const int add = chunk_base[blockIdx.x];
    const long long base = (long long)blockIdx.x * SC_CHUNK + (long long)threadIdx.x * 8;
#pragma unroll
    for (int q = 0; q < 8; ++q) if (base + q < n) out[base + q] += add;
}

// scratch ints needed by scan_exclusive for n elements (all levels of chunk sums)
static inline size_t scan_scratch_ints(long long n)
{
    size_t total = 0;
    while (n > SC_CHUNK) { n = (n + SC_CHUNK - 1) / SC_CHUNK; total += (size_t)n + 8; }
    return total + 8;
}
// out[i] = sum of in[0..i) (in == out allowed).  Returns the kernels launched.
static inline int scan_exclusive(const int* in, int* out, long long n, int* scratch, cudaStream_t s)
{
    if (n <= 0) return 0;
    const long long nchunks = (n + SC_CHUNK - 1) / SC_CHUNK;
    if (nchunks == 1) { k_scan_chunks<<<1, 256, 0, s>>>(in, out, nullptr, n); return 1; }
    k_scan_chunks<<<(unsigned)nchunks, 256, 0, s>>>(in, out, scratch, n);
    int launches = 1 + scan_exclusive(scratch, scratch, nchunks, scratch + nchunks + 8, s);
    k_scan_add<<<(unsigned)nchunks, 256, 0, s>>>(out, scratch, n);
    return launches + 1;
}

// hist ints needed for n elements
static inline size_t sort_hist_ints(long long n) { return (size_t)((n + RS_TILE - 1) / RS_TILE) * RS_BINS; }

// Sorts the pairs by the key bits covered by `shifts` (8-bit digits, least significant first), stable.  keys[0] / vals[0]
// hold the input; returns the index (0 or 1) of the buffer pair holding the result.  *launches += kernels launched.
static inline int sort_pairs(u64* keys[2], int* vals[2], long long n, const int* shifts, int n_shifts, int* hist, int* scan_scratch, cudaStream_t s, int* launches)
{
    const int ntiles = (int)((n + RS_TILE - 1) / RS_TILE);
    int cur = 0;
    for (int p = 0; p < n_shifts; ++p) {
        k_rs_count<<<ntiles, RS_THREADS, 0, s>>>(keys[cur], n, shifts[p], hist, ntiles);
        const int ls = scan_exclusive(hist, hist, (long long)ntiles * RS_BINS, scan_scratch, s);
        k_rs_scatter<<<ntiles, RS_THREADS, 0, s>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n, shifts[p], hist, ntiles);
        if (launches) *launches += 2 + ls;
        cur ^= 1;
    }
    return cur;
}

}  // namespace llsort
