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

__device__ __forceinline__ u64 shfl_u64(u64 v, int src)
{
    int lo = __shfl_sync(LL_FULL_MASK, (int)(unsigned)v, src);
    int hi = __shfl_sync(LL_FULL_MASK, (int)(unsigned)(v >> 32), src);
    return ((u64)(unsigned)hi << 32) | (unsigned)lo;
}
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

// ---- block-wide bitonic sort of 64-bit keys in shared memory --------------------------------------------------
// Sorts every aligned segment of SEG keys (power of two >= 64, N a multiple of SEG) ascending.  Stages whose
// partner distance is < 64 run in registers: a warp owns 64 consecutive keys (2 per lane) and exchanges them with
// shuffles, so shared memory is touched once per "register phase" instead of once per stage (10 passes instead of
// 45 for SEG = 512).  All threads of the block must call it; ends with a barrier.
__device__ __forceinline__ void bitonic_reg_stages(u64& k0, u64& k1, int ibase, int kk, int jmax, int lane)
{
    // element indices: i0 = ibase + lane (slot 0), i1 = i0 + 32 (slot 1); stages j = jmax, jmax/2, ..., 1
    for (int j = jmax; j > 0; j >>= 1) {
        if (j == 32) {
            const bool asc = ((ibase + lane) & kk) == 0;
            if ((k0 > k1) == asc) { const u64 t = k0; k0 = k1; k1 = t; }
        } else {
            const bool lower = (lane & j) == 0;
            const u64 o0 = shfl_xor_u64(k0, j), o1 = shfl_xor_u64(k1, j);
            const bool asc0 = ((ibase + lane) & kk) == 0, asc1 = ((ibase + 32 + lane) & kk) == 0;
            k0 = (lower == asc0) ? (k0 < o0 ? k0 : o0) : (k0 < o0 ? o0 : k0);
            k1 = (lower == asc1) ? (k1 < o1 ? k1 : o1) : (k1 < o1 ? o1 : k1);
        }
    }
}
__device__ __forceinline__ void block_bitonic_sort_u64(u64* keys, int N, int SEG)
{
    const int lane = lane_id(), w = warp_id(), nw = blockDim.x >> 5, nchunk = N >> 6;
    // phase 0: every 64-chunk fully sorted (k = 2..64) in registers, direction taken from the global network
    for (int c = w; c < nchunk; c += nw) {
        const int ibase = (c << 6) & (SEG - 1);
        u64 k0 = keys[(c << 6) + lane], k1 = keys[(c << 6) + 32 + lane];
        for (int kk = 2; kk <= 64; kk <<= 1) bitonic_reg_stages(k0, k1, ibase, SEG == 64 && kk == 64 ? 0 : kk, kk >> 1, lane);
        keys[(c << 6) + lane] = k0;
        keys[(c << 6) + 32 + lane] = k1;
    }
    __syncthreads();
    for (int kk = 128; kk <= SEG; kk <<= 1) {
        for (int j = kk >> 1; j >= 64; j >>= 1) {
            for (int t = threadIdx.x; t < (N >> 1); t += blockDim.x) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const u64 a = keys[i], b = keys[i | j];
                const bool asc = ((i & (SEG - 1)) & kk) == 0;
                if ((a > b) == asc) { keys[i] = b; keys[i | j] = a; }
            }
            __syncthreads();
        }
        for (int c = w; c < nchunk; c += nw) {
            const int ibase = (c << 6) & (SEG - 1);
            u64 k0 = keys[(c << 6) + lane], k1 = keys[(c << 6) + 32 + lane];
            bitonic_reg_stages(k0, k1, ibase, kk == SEG ? 0 : kk, 32, lane);
            keys[(c << 6) + lane] = k0;
            keys[(c << 6) + 32 + lane] = k1;
        }
        __syncthreads();
    }
}

// ring x azimuth-bin index: bin of a point / query; NB = bins per ring (power of two)
__device__ __forceinline__ int azimuth_bin(float x, float y, int NB)
{
    const float phi = atan2f(y, x);  // any deterministic function: build and query use the same one
    int b = (int)floorf((phi + 3.14159265f) * ((float)NB * 0.15915494f));
    return b < 0 ? 0 : (b >= NB ? NB - 1 : b);
}
