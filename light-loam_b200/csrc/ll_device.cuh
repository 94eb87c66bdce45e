// Device helpers shared by the kernels.  The whole library is compiled with -fmad=false so that plain
// fp32 / fp64 expressions keep the reference's x86-64 (no-FMA) rounding; fma() is used explicitly only
// where parity does not depend on it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef unsigned long long u64;

#define LL_FULL_MASK 0xffffffffu
#define LL_PI 3.14159265358979323846  // M_PI

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ u64 shfl_xor_u64(u64 v, int m)
{
    int lo = __shfl_xor_sync(LL_FULL_MASK, (int)(unsigned)v, m);
    int hi = __shfl_xor_sync(LL_FULL_MASK, (int)(unsigned)(v >> 32), m);
    return ((u64)(unsigned)hi << 32) | (unsigned)lo;
}
// warp-wide minimum of 64-bit keys with the hardware reduction (REDUX): first the high words, then the low words
// among the lanes that hold the minimal high word
__device__ __forceinline__ u64 warp_min_u64(u64 v)
{
    const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
    const unsigned mhi = __reduce_min_sync(LL_FULL_MASK, hi);
    const unsigned mlo = __reduce_min_sync(LL_FULL_MASK, hi == mhi ? lo : 0xFFFFFFFFu);
    return ((u64)mhi << 32) | mlo;
}
__device__ __forceinline__ double shfl_xor_f64(double v, int m)
{
    return __longlong_as_double((long long)shfl_xor_u64((u64)__double_as_longlong(v), m));
}
__device__ __forceinline__ double shfl_down_f64(double v, int d)
{
    int lo = __shfl_down_sync(LL_FULL_MASK, __double2loint(v), d);
    int hi = __shfl_down_sync(LL_FULL_MASK, __double2hiint(v), d);
    return __hiloint2double(hi, lo);
}

// fp32 squared distance in the reference's order ((dx*dx)+(dy*dy))+(dz*dz), no FMA (-fmad=false).
__device__ __forceinline__ float sqdist3(float ax, float ay, float az, float bx, float by, float bz)
{
    const float dx = ax - bx, dy = ay - by, dz = az - bz;
    return dx * dx + dy * dy + dz * dz;
}

// Eigen QuaternionBase::_transformVector in fp64: uv = 2 (u x v); v + w uv + u x uv  (q = x,y,z,w).
__device__ __forceinline__ void quat_rotate(const double q[4], double vx, double vy, double vz, double& ox, double& oy, double& oz)
{
    double ux = q[1] * vz - q[2] * vy, uy = q[2] * vx - q[0] * vz, uz = q[0] * vy - q[1] * vx;
    ux = ux + ux; uy = uy + uy; uz = uz + uz;
    const double cx = q[1] * uz - q[2] * uy, cy = q[2] * ux - q[0] * uz, cz = q[0] * uy - q[1] * ux;
    ox = (vx + q[3] * ux) + cx;
    oy = (vy + q[3] * uy) + cy;
    oz = (vz + q[3] * uz) + cz;
}
// Eigen quat_product a * b (Hamilton), x,y,z,w storage.
__device__ __forceinline__ void quat_mul(const double a[4], const double b[4], double o[4])
{
    const double x = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
    const double y = a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2];
    const double z = a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0];
    const double w = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
    o[0] = x; o[1] = y; o[2] = z; o[3] = w;
}

// Exclusive scan of one int per thread over a block of NT threads (NT multiple of 32, <= 1024).
// ws must hold 33 ints. Returns the exclusive prefix; *total receives the block sum.
__device__ __forceinline__ int block_exclusive_scan(int v, int* ws, int* total)
{
    const int lane = lane_id(), w = warp_id();
    int incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(LL_FULL_MASK, incl, d); if (lane >= d) incl += o; }
    __syncthreads();  // protect ws from a previous use
    if (lane == 31) ws[w] = incl;
    __syncthreads();
    if (w == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        int s = lane < nw ? ws[lane] : 0;
        int si = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(LL_FULL_MASK, si, d); if (lane >= d) si += o; }
        ws[lane] = si - s;
        if (lane == 31) ws[32] = si;
    }
    __syncthreads();
    const int r = ws[w] + incl - v;
    if (total) *total = ws[32];
    return r;
}

// hashed uniform grid: bucket of integer cell (ix,iy,iz), T = power of two
__device__ __forceinline__ int cell_bucket(int ix, int iy, int iz, int Tmask)
{
    const unsigned h = ((unsigned)ix * 73856093u) ^ ((unsigned)iy * 19349663u) ^ ((unsigned)iz * 83492791u);
    return (int)((h ^ (h >> 15)) & (unsigned)Tmask);
}

// ---- block-wide ascending sort of NS keys, 64- or 32-bit (shared memory <-> registers) ---------------------------
// All-ascending form of the bitonic network: the first stage of the merge of width kk pairs element e with its mirror
// image e ^ (kk - 1), the following ones e with e ^ j, and the lower index always keeps the smaller key - no stage
// needs a direction.  Thread t < NS / KPT holds elements t * KPT .. t * KPT + KPT - 1 in registers (blocked layout):
// distances below KPT are register-to-register, distances below 32 * KPT are shuffles inside the warp, and only the
// few largest distances of the last merges go through shared memory (with a block barrier each).  Fully unrolled.
// Every thread of the block must call it (barriers); on return keys[0 .. NS) is sorted and visible to the block.
__device__ __forceinline__ void cex_key(u64& lo, u64& hi)
{
    const u64 a = lo, c = hi;
    const bool sw = c < a;
    lo = sw ? c : a;
    hi = sw ? a : c;
}
__device__ __forceinline__ void cex_key(unsigned& lo, unsigned& hi)
{
    const unsigned a = lo, c = hi;
    lo = min(a, c);
    hi = max(a, c);
}
__device__ __forceinline__ u64 shfl_xor_key(u64 v, int m) { return shfl_xor_u64(v, m); }
__device__ __forceinline__ unsigned shfl_xor_key(unsigned v, int m) { return __shfl_xor_sync(LL_FULL_MASK, v, m); }
// what the element keeps after meeting `o`: the smaller key when it is the lower index of the pair, else the larger
__device__ __forceinline__ u64 keep_key(bool lower, u64 k, u64 o) { return ((o < k) == lower) ? o : k; }
__device__ __forceinline__ unsigned keep_key(bool lower, unsigned k, unsigned o) { return lower ? min(k, o) : max(k, o); }
template <typename K, int NS, int KPT>
__device__ __forceinline__ void block_sort_asc(K* keys)
{
    constexpr int NT = NS / KPT;          // threads holding keys
    constexpr int WSPAN = 32 * KPT;       // elements covered by one warp
    const int t = threadIdx.x, lane = t & 31;
    K k[KPT];
    if (t < NT) {
#pragma unroll
        for (int q = 0; q < KPT; ++q) k[q] = keys[t * KPT + q];
    }
#pragma unroll
    for (int kk = 2; kk <= NS; kk <<= 1) {
        if (kk > WSPAN) {
            // the stages of this merge that reach across warps run on the shared array
            if (t < NT) {
#pragma unroll
                for (int q = 0; q < KPT; ++q) keys[t * KPT + q] = k[q];
            }
            __syncthreads();
            for (int p = t; p < NS / 2; p += blockDim.x) {   // mirror stage
                const int blk = p / (kk / 2), off = p % (kk / 2);
                const int i = blk * kk + off, pr = blk * kk + (kk - 1 - off);
                K a = keys[i], c = keys[pr];
                cex_key(a, c);
                keys[i] = a; keys[pr] = c;
            }
            __syncthreads();
#pragma unroll
            for (int j = kk >> 2; j >= WSPAN; j >>= 1) {
                for (int p = t; p < NS / 2; p += blockDim.x) {
                    const int i = ((p & ~(j - 1)) << 1) | (p & (j - 1));
                    K a = keys[i], c = keys[i | j];
                    cex_key(a, c);
                    keys[i] = a; keys[i | j] = c;
                }
                __syncthreads();
            }
            if (t < NT) {
#pragma unroll
                for (int q = 0; q < KPT; ++q) k[q] = keys[t * KPT + q];
            }
            __syncthreads();   // the next round's stores must not overtake these loads
        }
        if (t < NT) {
            // mirror stage when it stays inside the warp
            if (kk <= KPT) {
#pragma unroll
                for (int q = 0; q < KPT; ++q) {
                    const int pq = q ^ (kk - 1);
                    if (q < pq) cex_key(k[q], k[pq]);
                }
            } else if (kk <= WSPAN) {
                const int m = kk / KPT - 1;
                const bool lower = (lane & (kk / (2 * KPT))) == 0;
                K o[KPT];
#pragma unroll
                for (int q = 0; q < KPT; ++q) o[q] = shfl_xor_key(k[KPT - 1 - q], m);
#pragma unroll
                for (int q = 0; q < KPT; ++q) k[q] = keep_key(lower, k[q], o[q]);
            }
            // e ^ j stages inside the warp
#pragma unroll
            for (int j = (kk >> 2) < WSPAN ? (kk >> 2) : (WSPAN >> 1); j > 0; j >>= 1) {
                if (j < KPT) {
#pragma unroll
                    for (int q = 0; q < KPT; ++q)
                        if ((q & j) == 0) cex_key(k[q], k[q | j]);
                } else {
                    const int m = j / KPT;
                    const bool lower = (lane & m) == 0;
#pragma unroll
                    for (int q = 0; q < KPT; ++q) k[q] = keep_key(lower, k[q], shfl_xor_key(k[q], m));
                }
            }
        }
    }
    if (t < NT) {
#pragma unroll
        for (int q = 0; q < KPT; ++q) keys[t * KPT + q] = k[q];
    }
    __syncthreads();
}
template <int NS, int KPT>
__device__ __forceinline__ void block_sort_u64_asc(u64* keys) { block_sort_asc<u64, NS, KPT>(keys); }

// ring x azimuth-bin index: bin of a point / query; NB = bins per ring (power of two)
__device__ __forceinline__ int azimuth_bin(float x, float y, int NB)
{
    const float phi = atan2f(y, x);  // any deterministic function: build and query use the same one
    int b = (int)floorf((phi + 3.14159265f) * ((float)NB * 0.15915494f));
    return b < 0 ? 0 : (b >= NB ? NB - 1 : b);
}

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier ------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// several copies behind ONE expect_tx: announce the total first, then issue the copies without arriving again
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_copy(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
