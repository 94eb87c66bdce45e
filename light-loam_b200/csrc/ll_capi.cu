// C ABI of liblightloam_b200 (include/lightloam_b200.h): context, staging, and the reference-facing calls.
// No CPU fallback anywhere: every entry point runs the CUDA kernels or returns an error code.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <stddef.h>
#include <string.h>

#include <new>
#include <vector>

#include "ll_ctx.h"

size_t ll_feature_smem_bytes(int SCAP);

namespace {

struct ScanHdr { int n_raw, stride_words; unsigned off_lo, off_hi; };   // off = word offset of the scan inside the staging slab

__global__ void k_set_scan_hdr(LaneState* lane, const ScanHdr* hdr, const uint32_t* raw, int n_lanes)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n_lanes) {
        lane[b].raw = raw + (((size_t)hdr[b].off_hi << 32) | hdr[b].off_lo);
        lane[b].n_raw = hdr[b].n_raw;
        lane[b].stride_words = hdr[b].stride_words;
        lane[b].err = 0;
    }
}

__global__ void k_set_pool_hdr(LaneState* lane, const int* ids, const int* pool_n, const uint32_t* pool, size_t scan_words, int n_lanes)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < n_lanes) {
        lane[b].raw = pool + (size_t)ids[b] * scan_words;
        lane[b].n_raw = pool_n[ids[b]];
        lane[b].stride_words = 4;
        lane[b].err = 0;
    }
}

// float4 x,y,z,intensity -> pcl::PointXYZI as pcl::toROSMsg lays it out in sensor_msgs/PointCloud2::data (point_step 32:
// x@0 y@4 z@8 data[3]@12 = 1.0f intensity@16, padding zeroed) - SR:382-410, LO:899-913.  Two 16-byte stores per point, coalesced.
__global__ void k_pack_pointcloud2(const float4* __restrict__ src, float4* __restrict__ dst, const int* n_dev, int n_host)
{
    const int n = n_dev ? *n_dev : n_host;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 p = src[i];
        dst[2 * i] = make_float4(p.x, p.y, p.z, 1.0f);      // PCL_ADD_POINT4D: data[3] = 1.0f (set by the PointXYZI constructor)
        dst[2 * i + 1] = make_float4(p.w, 0.f, 0.f, 0.f);
    }
}

__global__ void k_init_lanes(LaneState* lane, int n_lanes)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_lanes) return;
    LaneState z;
    memset(&z, 0, sizeof(z));
    z.para_q[3] = 1.0;  // LO:61
    z.q_w[3] = 1.0;     // LO:57
    z.map_par[3] = 1.0; // LM:81
    z.q_wmap_wodom[3] = 1.0;  // LM:87
    z.last_slot = 1;
    z.cen[0] = 10; z.cen[1] = 10; z.cen[2] = 5;  // LM:42-44
    lane[b] = z;
}

// odometry fed from host clouds: emulate what feature extraction leaves behind in the lane
__global__ void k_set_feature_counts(LaneState* lane, int ns, int nls, int nf, int nlf)
{
    LaneState& L = lane[0];
    L.cur = L.last_slot ^ 1;
    L.n_sharp = ns; L.n_less_sharp = nls; L.n_flat = nf; L.n_less_flat = nlf;
    L.err = 0;
}


template <typename T>
cudaError_t dalloc(T*& p, size_t n) { return cudaMalloc((void**)&p, n * sizeof(T)); }

// smallest fp32 t >= 0 with expf(-t) < 0.96f under this host's libm: LO:239-242 decides votes with
// `std::exp(-(gap*gap)) < 0.96f`; the device compares gap*gap >= t instead of evaluating exp.
float calibrate_vote_threshold()
{
    auto bits = [](float f) { uint32_t u; memcpy(&u, &f, 4); return u; };
    auto flt = [](uint32_t u) { float f; memcpy(&f, &u, 4); return f; };
    uint32_t lo = bits(0.03f);  // expf(-0.03) = 0.970 >= 0.96
    uint32_t hi = bits(0.05f);  // expf(-0.05) = 0.951 <  0.96
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (expf(-flt(mid)) < 0.96f) hi = mid; else lo = mid;
    }
    return flt(hi);
}

// the same for laserMapping.cpp's copy (LM:918-924): float score compared with the double literal 0.95
float calibrate_vote_threshold95()
{
    auto bits = [](float f) { uint32_t u; memcpy(&u, &f, 4); return u; };
    auto flt = [](uint32_t u) { float f; memcpy(&f, &u, 4); return f; };
    uint32_t lo = bits(0.04f);  // expf(-0.04) = 0.961 >= 0.95
    uint32_t hi = bits(0.06f);  // expf(-0.06) = 0.942 <  0.95
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if ((double)expf(-flt(mid)) < 0.95) hi = mid; else lo = mid;
    }
    return flt(hi);
}

int copy_view_to_device(ll_ctx* c, const ll_cloud_view& v, float4* dst, int cap)
{
    if (v.n < 0 || v.n > cap) return LL_E_CAPACITY;
    if (v.n == 0) return LL_OK;
    if (!v.data || v.stride_bytes < 12 || (v.stride_bytes & 3)) return LL_E_INVAL;
    if (v.stride_bytes == 16) {
        LL_CUDA_CHECK(c, cudaMemcpyAsync(dst, v.data, (size_t)v.n * 16, cudaMemcpyHostToDevice, c->stream));
    } else if (v.stride_bytes >= 32) {
        // pcl::PointXYZI as it travels in sensor_msgs/PointCloud2 (point_step 32): x,y,z at 0, intensity at byte 16
        LL_CUDA_CHECK(c, cudaMemcpy2DAsync(dst, 16, v.data, v.stride_bytes, 12, v.n, cudaMemcpyHostToDevice, c->stream));
        LL_CUDA_CHECK(c, cudaMemcpy2DAsync(reinterpret_cast<char*>(dst) + 12, 16, reinterpret_cast<const char*>(v.data) + 16, v.stride_bytes, 4, v.n,
                                           cudaMemcpyHostToDevice, c->stream));
    } else {
        LL_CUDA_CHECK(c, cudaMemcpy2DAsync(dst, 16, v.data, v.stride_bytes, 16 <= v.stride_bytes ? 16 : 12, v.n, cudaMemcpyHostToDevice, c->stream));
    }
    return LL_OK;
}

int copy_out(ll_ctx* c, ll_cloud_out* o, const float4* src, int n)
{
    if (!o) return LL_OK;
    o->n = 0;
    if (n > o->cap) return LL_E_CAPACITY;
    if (n > 0 && o->xyzi) LL_CUDA_CHECK(c, cudaMemcpyAsync(o->xyzi, src, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream));
    o->n = n;
    return LL_OK;
}

}  // namespace

extern "C" {

void ll_default_config(ll_config* cfg, int scan_line)
{
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->scan_line = scan_line;
    // launch/aloam_velodyne_HDL_64.launch:2-12 ; launch/aloam_velodyne_VLP_16.launch:3-13 (HDL-32 identical)
    cfg->minimum_range = scan_line == 64 ? 5.0f : 0.3f;
    cfg->lower_bound = -24.9f;  // SR:439
    cfg->up_bound = 2.0f;       // SR:440
    cfg->line_res = scan_line == 64 ? 0.4f : 0.2f;
    cfg->plane_res = scan_line == 64 ? 0.8f : 0.4f;
    cfg->skip_frame = 1;
    cfg->graph_from_frame = 5;
    cfg->device = 0;
    cfg->batch = 1;
    cfg->max_points = scan_line == 64 ? 131072 : (scan_line == 32 ? 73728 : 32768);
    // Ring capacity: the 64-line formula (SR:162) bins real HDL-64 elevations uniformly although the beams are not, so a
    // scanID can collect two lasers (3.5-4k points); a VLP-16 at 5 Hz also has ~3.6k points per ring.  6155 = 6 x 1024 + 11
    // enables the wide-sector kernels next to the 512-key ones (rings of up to 3083 points keep the fast path).
    cfg->max_ring_points = 6155;
    cfg->map_capacity = 1 << 20;
    cfg->enable_mapping = 0;
}

const char* ll_strerror(int code)
{
    switch (code) {
        case LL_OK: return "ok";
        case LL_E_INVAL: return "invalid argument";
        case LL_E_CAPACITY: return "input exceeds configured capacity";
        case LL_E_CUDA: return "CUDA error";
        case LL_E_NCCL: return "NCCL error";
        case LL_E_EMPTY: return "scan has no valid point";
        case LL_W_FEW_CORRESPONDENCES: return "warning: too few correspondences / map too small";
        default: return "unknown error";
    }
}

const char* ll_last_error(const ll_ctx* ctx) { return ctx ? ctx->last_error.c_str() : ""; }

void ll_destroy(ll_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->dev);
    if (c->stream) cudaStreamSynchronize(c->stream);
    ll_map_free(c);
    for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
    for (cudaEvent_t e : c->prof_ev) cudaEventDestroy(e);
    for (int k = 0; k < 2; ++k) {
        if (c->ev_staged[k]) cudaEventDestroy(c->ev_staged[k]);
        if (c->ev_raw_free[k]) cudaEventDestroy(c->ev_raw_free[k]);
        if (c->ev_done[k]) cudaEventDestroy(c->ev_done[k]);
    }
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->d_raw2) cudaFree(c->d_raw2);
    if (c->d_hdr2) cudaFree(c->d_hdr2);
    if (c->h_hdr2) cudaFreeHost(c->h_hdr2);
    if (c->h_pose2) cudaFreeHost(c->h_pose2);
    if (c->h_ids) cudaFreeHost(c->h_ids);
    if (c->h_status) cudaFreeHost(c->h_status);
    free(c->last_status);
    void* ptrs[] = {c->d_odom_comm, c->d_odom_seq[0], c->d_odom_seq[1], c->d_wide_list, c->d_wide_n, c->d_status, c->d_pc2, c->d_qa, c->d_qb, c->d_qstart, c->d_pool, c->d_pool_n, c->d_ids, c->d_lane, c->d_pose, c->d_hdr, c->d_raw, c->d_ring8, c->d_rank8, c->d_tile_hist, c->d_full, c->d_curv, c->d_label, c->d_sorted16, c->d_brk, c->d_lf_tmp,
                    c->d_ring_lists, c->d_ring_counts, c->d_sharp, c->d_flat, c->d_sharp_idx, c->d_lsharp_idx, c->d_flat_idx,
                    c->d_lsharp[0], c->d_lsharp[1], c->d_lflat[0], c->d_lflat[1], c->d_ebound[0], c->d_ebound[1], c->d_bands[0], c->d_bands[1], c->a_corner.start, c->a_corner.cursor, c->a_corner.sorted,
                    c->a_corner.partial, c->a_surf.start, c->a_surf.cursor, c->a_surf.sorted, c->a_surf.partial, c->d_corner_assoc, c->d_plane_assoc, c->d_blocks, c->d_assoc_queue, c->d_assoc_queue_n, c->d_vote_src, c->d_vote_tgt};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (c->h_lane) cudaFreeHost(c->h_lane);
    if (c->h_pose) cudaFreeHost(c->h_pose);
    if (c->h_hdr) cudaFreeHost(c->h_hdr);
    for (cudaEvent_t e : c->ev) if (e) cudaEventDestroy(e);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int ll_create(const ll_config* cfg, ll_ctx** out)
{
    if (!cfg || !out) return LL_E_INVAL;
    *out = nullptr;
    if (cfg->scan_line != 16 && cfg->scan_line != 32 && cfg->scan_line != 64) return LL_E_INVAL;  // SR:447-451
    if (cfg->batch < 1 || cfg->batch > 4096 || cfg->max_points < 256 || cfg->max_ring_points < 17 || cfg->max_ring_points > 6 * 1024 + 11)
        return LL_E_INVAL;
    ll_ctx* c = new (std::nothrow) ll_ctx();
    if (!c) return LL_E_INVAL;
    c->cfg = *cfg;
    c->B = cfg->batch;
    c->R = cfg->scan_line;
    c->Nmax = (cfg->max_points + LL_TILE - 1) / LL_TILE * LL_TILE;
    c->NT = c->Nmax / LL_TILE;
    c->SCAP = cfg->max_ring_points <= 6 * 512 + 11 ? 512 : 1024;
    c->RCAP = 6 * c->SCAP + 16;
    c->dev = cfg->device;
    c->vote_t_min = calibrate_vote_threshold();
    c->vote_t95 = calibrate_vote_threshold95();
#define CK(expr)                                                                     \
    do {                                                                             \
        cudaError_t e__ = (expr);                                                    \
        if (e__ != cudaSuccess) {                                                    \
            fprintf(stderr, "lightloam_b200: %s: %s\n", #expr, cudaGetErrorString(e__)); \
            ll_destroy(c);                                                           \
            return LL_E_CUDA;                                                        \
        }                                                                            \
    } while (0)
    CK(cudaSetDevice(c->dev));
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    for (auto& e : c->ev) CK(cudaEventCreate(&e));
    const size_t B = c->B, N = c->Nmax, R = c->R;
    CK(dalloc(c->d_lane, B));
    CK(cudaHostAlloc((void**)&c->h_lane, sizeof(LaneState) * B, cudaHostAllocDefault));
    CK(cudaHostAlloc((void**)&c->h_pose, sizeof(double) * 14 * B, cudaHostAllocDefault));
    CK(dalloc(c->d_pose, B * 14));
    CK(cudaHostAlloc((void**)&c->h_hdr, sizeof(int) * 4 * B, cudaHostAllocDefault));
    CK(dalloc(c->d_hdr, B * 4));
    CK(dalloc(c->d_wide_list, B * R));
    CK(dalloc(c->d_wide_n, 1));
    CK(cudaMemsetAsync(c->d_wide_n, 0, sizeof(int), c->stream));
    CK(dalloc(c->d_status, B));
    CK(cudaMemsetAsync(c->d_status, 0, sizeof(int) * B, c->stream));
    CK(cudaHostAlloc((void**)&c->h_status, sizeof(int) * 3 * B, cudaHostAllocDefault));
    memset(c->h_status, 0, sizeof(int) * 3 * B);
    c->last_status = (int*)calloc(B, sizeof(int));
    CK(dalloc(c->d_raw, B * N * 8));
    CK(dalloc(c->d_ring8, B * N));
    CK(dalloc(c->d_rank8, B * N));
    CK(dalloc(c->d_tile_hist, B * c->NT * R));
    CK(dalloc(c->d_full, B * N));
    CK(dalloc(c->d_curv, B * N));
    CK(dalloc(c->d_label, B * N));
    CK(dalloc(c->d_sorted16, B * N));
    CK(dalloc(c->d_brk, B * R * (size_t)((c->RCAP + 31) / 32)));
    CK(dalloc(c->d_lf_tmp, B * N));
    CK(dalloc(c->d_ring_lists, B * R * (LL_SHARP_PER_RING + LL_LSHARP_PER_RING + LL_FLAT_PER_RING)));
    CK(dalloc(c->d_ring_counts, B * R * 4));
    CK(dalloc(c->d_sharp, B * R * LL_SHARP_PER_RING));
    CK(dalloc(c->d_flat, B * R * LL_FLAT_PER_RING));
    CK(dalloc(c->d_sharp_idx, B * R * LL_SHARP_PER_RING));
    CK(dalloc(c->d_lsharp_idx, B * R * LL_LSHARP_PER_RING));
    CK(dalloc(c->d_flat_idx, B * R * LL_FLAT_PER_RING));
    for (int k = 0; k < 2; ++k) {
        CK(dalloc(c->d_lsharp[k], B * R * LL_LSHARP_PER_RING));
        CK(dalloc(c->d_lflat[k], B * N));
    }
    // polar index over the previous frame's less-sharp / less-flat clouds: bucket = azimuth bin * R + ring.  It serves
    // both the exact 1-NN (kdtree*Last, LO:494 / LO:656) and the ring-window search of LO:504-553 / LO:668-721.
    c->az_bins_corner = 64; c->az_bins_surf = 256;
    if (const char* e = getenv("LL_AZ_CORNER")) { const int v = atoi(e); if (v >= 8 && v <= 4096 && (v & (v - 1)) == 0) c->az_bins_corner = v; }
    if (const char* e = getenv("LL_AZ_SURF")) { const int v = atoi(e); if (v >= 8 && v <= 4096 && (v & (v - 1)) == 0) c->az_bins_surf = v; }
    while ((int)R * c->az_bins_corner < 2048) c->az_bins_corner <<= 1;
    while ((int)R * c->az_bins_surf < 2048) c->az_bins_surf <<= 1;
    c->a_corner.T = (int)R * c->az_bins_corner; c->a_corner.cap = (int)R * LL_LSHARP_PER_RING;
    c->a_surf.T = (int)R * c->az_bins_surf; c->a_surf.cap = (int)N;
    for (KnnGrid* g : {&c->a_corner, &c->a_surf}) {
        CK(dalloc(g->start, B * (size_t)(g->T + 1)));
        CK(dalloc(g->cursor, B * (size_t)g->T));
        CK(dalloc(g->partial, B * (size_t)(g->T / 2048 + 1)));
        CK(dalloc(g->sorted, B * (size_t)g->cap));
    }
    for (int k = 0; k < 2; ++k) { CK(dalloc(c->d_ebound[k], B * R * 2)); CK(dalloc(c->d_bands[k], B * (2 * LL_MAX_RINGS + 4))); }
    CK(dalloc(c->d_corner_assoc, B * R * LL_SHARP_PER_RING * 2));
    CK(dalloc(c->d_plane_assoc, B * R * LL_FLAT_PER_RING * 4));
    CK(dalloc(c->d_vote_src, B * R * LL_FLAT_PER_RING));
    CK(dalloc(c->d_vote_tgt, B * R * LL_FLAT_PER_RING));
    c->assoc_queue_cap = (int)(B * R * (LL_SHARP_PER_RING + LL_FLAT_PER_RING));
    CK(dalloc(c->d_assoc_queue, (size_t)c->assoc_queue_cap));
    CK(dalloc(c->d_assoc_queue_n, 16));
    CK(cudaMemsetAsync(c->d_assoc_queue_n, 0, sizeof(int) * 16, c->stream));
    CK(dalloc(c->d_qa, B * R * (LL_SHARP_PER_RING + LL_FLAT_PER_RING)));
    CK(dalloc(c->d_qb, B * R * (LL_SHARP_PER_RING + LL_FLAT_PER_RING)));
    CK(dalloc(c->d_qstart, B * (size_t)c->qstart_stride));
    {   // mailbox of the split odometry solve: [lane][parity][128 contributions][32 doubles] + [lane][128] flags (ll_solve.cuh)
        c->odom_comm_mbox_bytes = sizeof(double) * B * 2 * 128 * 32;
        const size_t total = c->odom_comm_mbox_bytes + sizeof(unsigned long long) * B * 128;
        CK(cudaMalloc(&c->d_odom_comm, total));
        CK(cudaMemsetAsync(c->d_odom_comm, 0, total, c->stream));
        for (int k = 0; k < 2; ++k) { CK(dalloc(c->d_odom_seq[k], B)); CK(cudaMemsetAsync(c->d_odom_seq[k], 0, sizeof(unsigned long long) * B, c->stream)); }
    }
    c->nblk_cap = (int)R * (LL_SHARP_PER_RING + LL_FLAT_PER_RING);
    CK(dalloc(c->d_blocks, B * (size_t)LL_BLOCK_DOUBLES * c->nblk_cap));
    CK(cudaMemsetAsync(c->d_raw, 0, sizeof(uint32_t) * B * N * 8, c->stream));
    k_init_lanes<<<(c->B + 63) / 64, 64, 0, c->stream>>>(c->d_lane, c->B);
    CK(cudaGetLastError());
    if (cfg->enable_mapping) {
        const int rc = ll_map_alloc(c);
        if (rc != LL_OK) { ll_destroy(c); return rc; }
    }
    CK(cudaStreamSynchronize(c->stream));
#undef CK
    *out = c;
    return LL_OK;
}

int ll_reset(ll_ctx* c)
{
    if (!c) return LL_E_INVAL;
    LL_CUDA_CHECK(c, cudaSetDevice(c->dev));
    k_init_lanes<<<(c->B + 63) / 64, 64, 0, c->stream>>>(c->d_lane, c->B);
    if (c->map) ll_map_clear(c);  // empties the cube map in place: buffers (and the mailbox other ranks may hold) stay put
    LL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return LL_OK;
}

void* ll_cuda_stream(ll_ctx* c) { return c ? (void*)c->stream : nullptr; }
int ll_launch_count(const ll_ctx* c) { return c ? c->launches : 0; }

static int fetch_lanes(ll_ctx* c, int n)
{
    LL_CUDA_CHECK(c, cudaMemcpyAsync(c->h_lane, c->d_lane, sizeof(LaneState) * n, cudaMemcpyDeviceToHost, c->stream));
    LL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return LL_OK;
}

// Raw-scan records a kernel can read in place: x,y,z at words 0..2 of a record of 3..8 words.  Anything else (point_step
// above 32 or not a multiple of 4: velodyne XYZIRT 22 B, Ouster 48 B ...) is gathered to packed xyz by a strided copy,
// which is what pcl::fromROSMsg does on the host in the reference (SR:105-106).
static bool stride_in_place(int stride_bytes) { return stride_bytes >= 12 && stride_bytes <= 32 && (stride_bytes & 3) == 0; }
static int check_scan_view(const ll_ctx* c, const ll_cloud_view& v)
{
    if (!v.data || v.n < 1 || v.stride_bytes < 12) return LL_E_INVAL;
    if (v.n > c->Nmax) return LL_E_CAPACITY;
    return LL_OK;
}
// enqueues the copy of one scan into lane slot i of a staging slab and fills its header
static int stage_one(ll_ctx* c, const ll_cloud_view& v, uint32_t* slab, int i, ScanHdr* hdr, cudaStream_t st)
{
    const size_t off = (size_t)i * c->Nmax * 8;
    hdr[i].n_raw = v.n;
    hdr[i].off_lo = (unsigned)off; hdr[i].off_hi = (unsigned)(off >> 32);
    if (stride_in_place(v.stride_bytes)) {
        hdr[i].stride_words = v.stride_bytes / 4;
        LL_CUDA_CHECK(c, cudaMemcpyAsync(slab + off, v.data, (size_t)v.n * v.stride_bytes, cudaMemcpyHostToDevice, st));
    } else {
        hdr[i].stride_words = 3;
        LL_CUDA_CHECK(c, cudaMemcpy2DAsync(slab + off, 12, v.data, v.stride_bytes, 12, v.n, cudaMemcpyHostToDevice, st));
    }
    return LL_OK;
}

int ll_stage_scans(ll_ctx* c, int n_scans, const ll_cloud_view* scans)
{
    if (!c || !scans || n_scans < 1 || n_scans > c->B) return LL_E_INVAL;
    LL_CUDA_CHECK(c, cudaSetDevice(c->dev));
    if (c->sub_count > 0) return LL_E_INVAL;           // the staging slab belongs to an outstanding submission
    LL_CUDA_CHECK(c, cudaEventSynchronize(c->ev[4]));  // the previous staging must have consumed h_hdr
    ScanHdr* hdr = reinterpret_cast<ScanHdr*>(c->h_hdr);
    for (int i = 0; i < n_scans; ++i) { const int rc = check_scan_view(c, scans[i]); if (rc) return rc; }
    for (int i = 0; i < n_scans; ++i) { const int rc = stage_one(c, scans[i], c->d_raw, i, hdr, c->stream); if (rc) return rc; }
    LL_CUDA_CHECK(c, cudaMemcpyAsync(c->d_hdr, hdr, sizeof(ScanHdr) * n_scans, cudaMemcpyHostToDevice, c->stream));
    LL_CUDA_CHECK(c, cudaEventRecord(c->ev[4], c->stream));
    k_set_scan_hdr<<<(n_scans + 63) / 64, 64, 0, c->stream>>>(c->d_lane, reinterpret_cast<const ScanHdr*>(c->d_hdr), c->d_raw, n_scans);
    c->pre_launches = 1;
    LL_CUDA_CHECK(c, cudaGetLastError());
    return LL_OK;
}

// first non-zero lane status of a finished batch (kept for ll_get_lane_status)
static int publish_status(ll_ctx* c, const int* st, int n)
{
    int first = LL_OK;
    for (int i = 0; i < n; ++i) { c->last_status[i] = st[i]; if (first == LL_OK && st[i] < 0) first = st[i]; }
    return first;
}

// The kernel sequence of one step (SR:100-377 -> LO:425-896 [-> LM:1581-2168]) on the context stream.
static int launch_pipeline(ll_ctx* c, int n_scans, bool with_events)
{
    if (with_events) LL_CUDA_CHECK(c, cudaEventRecord(c->ev[0], c->stream));
    int rc = ll_launch_features(c, n_scans);
    if (rc) return rc;
    if (with_events) LL_CUDA_CHECK(c, cudaEventRecord(c->ev[1], c->stream));
    rc = ll_launch_odometry(c, n_scans);
    if (rc) return rc;
    if (with_events) LL_CUDA_CHECK(c, cudaEventRecord(c->ev[2], c->stream));
    if (c->cfg.enable_mapping) {
        rc = ll_launch_mapping(c, n_scans);
        if (rc) return rc;
    }
    if (with_events) LL_CUDA_CHECK(c, cudaEventRecord(c->ev[3], c->stream));
    return LL_OK;
}

// Latency path: with few lanes a step is ~30 short kernels and the launches cost as much as the work.  The sequence is
// captured once per (lane count, mailbox parity) into a CUDA graph and replayed; LL_GRAPH=0 disables, LL_GRAPH=1 forces it
// for any lane count.  With mapping the key also carries which cube-map buffer is current (the frames alternate).  Not used
// while profiling, nor when the mapping solve runs its collectives through the mailbox (several GPUs, LL_LM_PARTS).
static bool graph_wanted(const ll_ctx* c, int n_scans)
{
    if (c->prof || c->eager_calls < 2) return false;
    if (c->cfg.enable_mapping && ll_map_graph_state(c) < 0) return false;
    if (const char* e = getenv("LL_GRAPH")) return atoi(e) != 0;
    return n_scans <= 16;
}

int ll_process_staged(ll_ctx* c, int n_scans, double* poses_out)
{
    if (!c || n_scans < 1 || n_scans > c->B) return LL_E_INVAL;
    LL_CUDA_CHECK(c, cudaSetDevice(c->dev));
    if (c->prof) ll_prof_harvest(c);
    const int pre = c->pre_launches;   // the header kernel of ll_stage_scans / ll_process_pool belongs to this step
    c->pre_launches = 0;
    c->last_call_graph = false;
    if (graph_wanted(c, n_scans)) {
        const int map_state = c->cfg.enable_mapping ? ll_map_graph_state(c) : 0;
        const int key = (n_scans * 2 + c->odom_comm_flip) * 2 + map_state;
        auto it = c->graphs.find(key);
        if (it == c->graphs.end()) {
            cudaGraph_t g = nullptr;
            cudaGraphExec_t ge = nullptr;
            const int flip0 = c->odom_comm_flip;
            LL_CUDA_CHECK(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
            c->launches = 0;
            const int rc = launch_pipeline(c, n_scans, false);
            const cudaError_t e = cudaStreamEndCapture(c->stream, &g);
            c->odom_comm_flip = flip0;       // the capture only recorded: nothing ran yet
            if (c->cfg.enable_mapping) ll_map_graph_set_state(c, map_state);
            if (rc || e != cudaSuccess) { if (g) cudaGraphDestroy(g); c->last_error = std::string("graph capture: ") + cudaGetErrorString(e); return rc ? rc : LL_E_CUDA; }
            LL_CUDA_CHECK(c, cudaGraphInstantiate(&ge, g, 0));
            cudaGraphDestroy(g);
            c->graph_launches[key] = c->launches;
            it = c->graphs.emplace(key, ge).first;
        }
        LL_CUDA_CHECK(c, cudaGraphLaunch(it->second, c->stream));
        c->launches = pre + c->graph_launches[key];
        c->odom_comm_flip ^= c->graph_parity_step;   // three solves per step: the mailbox parity moves as in the eager calls before
        if (c->cfg.enable_mapping) ll_map_graph_set_state(c, map_state ^ 1);   // the frame wrote the other cube-map buffer
        c->last_call_graph = true;
    } else {
        c->launches = pre;
        const int flip0 = c->odom_comm_flip;
        const int rc = launch_pipeline(c, n_scans, true);
        if (rc) return rc;
        c->graph_parity_step = c->odom_comm_flip ^ flip0;
        c->eager_calls++;
    }
    if (poses_out) {
        LL_CUDA_CHECK(c, cudaMemcpyAsync(c->h_pose, c->d_pose, sizeof(double) * 14 * n_scans, cudaMemcpyDeviceToHost, c->stream));
        LL_CUDA_CHECK(c, cudaMemcpyAsync(c->h_status, c->d_status, sizeof(int) * n_scans, cudaMemcpyDeviceToHost, c->stream));
        LL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
        memcpy(poses_out, c->h_pose, sizeof(double) * 14 * n_scans);
        return publish_status(c, c->h_status, n_scans);
    }
    return LL_OK;
}

int ll_process_scans(ll_ctx* c, int n_scans, const ll_cloud_view* scans, double* poses_out)
{
    int rc = ll_stage_scans(c, n_scans, scans);
    if (rc) return rc;
    rc = ll_process_staged(c, n_scans, poses_out);
    if (rc) return rc;
    if (!poses_out) {
        LL_CUDA_CHECK(c, cudaMemcpyAsync(c->h_status, c->d_status, sizeof(int) * n_scans, cudaMemcpyDeviceToHost, c->stream));
        LL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
        return publish_status(c, c->h_status, n_scans);
    }
    return LL_OK;
}

// ---- scan pool: keep many scans resident in HBM, feed lanes by scan id ------------------------------------
int ll_pool_upload(ll_ctx* c, int n_scans, const ll_cloud_view* scans)
{
    if (!c || !scans || n_scans < 1) return LL_E_INVAL;
    LL_CUDA_CHECK(c, cudaSetDevice(c->dev));
    LL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    if (n_scans > c->pool_cap) {
        if (c->d_pool) cudaFree(c->d_pool);
        if (c->d_pool_n) cudaFree(c->d_pool_n);
        c->d_pool = nullptr; c->d_pool_n = nullptr; c->pool_cap = 0;
        LL_CUDA_CHECK(c, cudaMalloc((void**)&c->d_pool, sizeof(uint32_t) * 4 * (size_t)c->Nmax * n_scans));
        LL_CUDA_CHECK(c, cudaMalloc((void**)&c->d_pool_n, sizeof(int) * n_scans));
        c->pool_cap = n_scans;
    }
    if (!c->h_ids) {
        LL_CUDA_CHECK(c, cudaHostAlloc((void**)&c->h_ids, sizeof(int) * c->B, cudaHostAllocDefault));
        LL_CUDA_CHECK(c, cudaMalloc((void**)&c->d_ids, sizeof(int) * c->B));
    }
    std::vector<int> ns(n_scans);
    for (int i = 0; i < n_scans; ++i) {
        const ll_cloud_view& v = scans[i];
        if (!v.data || v.n < 1 || v.stride_bytes < 12) return LL_E_INVAL;
        if (v.n > c->Nmax) return LL_E_CAPACITY;
        ns[i] = v.n;
        uint32_t* dst = c->d_pool + (size_t)i * c->Nmax * 4;
        if (v.stride_bytes == 16)
            LL_CUDA_CHECK(c, cudaMemcpyAsync(dst, v.data, (size_t)v.n * 16, cudaMemcpyHostToDevice, c->stream));
        else
            LL_CUDA_CHECK(c, cudaMemcpy2DAsync(dst, 16, v.data, v.stride_bytes, v.stride_bytes < 16 ? 12 : 16, v.n, cudaMemcpyHostToDevice, c->stream));
    }
    LL_CUDA_CHECK(c, cudaMemcpyAsync(c->d_pool_n, ns.data(), sizeof(int) * n_scans, cudaMemcpyHostToDevice, c->stream));
    LL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    c->pool_n = n_scans;
    return LL_OK;
}

int ll_process_pool(ll_ctx* c, int n_lanes, const int* scan_ids, double* poses_out)
{
    if (!c || !scan_ids || n_lanes < 1 || n_lanes > c->B || !c->d_pool) return LL_E_INVAL;
    LL_CUDA_CHECK(c, cudaSetDevice(c->dev));
    LL_CUDA_CHECK(c, cudaEventSynchronize(c->ev[4]));
    for (int i = 0; i < n_lanes; ++i) {
        if (scan_ids[i] < 0 || scan_ids[i] >= c->pool_n) return LL_E_INVAL;
        c->h_ids[i] = scan_ids[i];
    }
    LL_CUDA_CHECK(c, cudaMemcpyAsync(c->d_ids, c->h_ids, sizeof(int) * n_lanes, cudaMemcpyHostToDevice, c->stream));
    LL_CUDA_CHECK(c, cudaEventRecord(c->ev[4], c->stream));
    k_set_pool_hdr<<<(n_lanes + 63) / 64, 64, 0, c->stream>>>(c->d_lane, c->d_ids, c->d_pool_n, c->d_pool, (size_t)c->Nmax * 4, n_lanes);
    c->pre_launches = 1;
    return ll_process_staged(c, n_lanes, poses_out);
}

// ---- per-kernel device timing ---------------------------------------------------------------------------------
int ll_profile_enable(ll_ctx* c, int on)
{
    if (!c) return LL_E_INVAL;
    if (c->prof) ll_prof_harvest(c);
    c->prof = on != 0;
    c->prof_acc.clear();
    return LL_OK;
}
// writes up to cap entries: names as a '\n'-separated list into names_buf; total ms and launches per name
int ll_profile_read(ll_ctx* c, char* names_buf, int buf_len, double* total_ms, int* launches, int cap)
{
    if (!c || !names_buf || !total_ms || !launches) return LL_E_INVAL;
    ll_prof_harvest(c);
    int k = 0;
    std::string names;
    for (const auto& kv : c->prof_acc) {
        if (k >= cap) break;
        names += kv.first;
        names += '\n';
        total_ms[k] = kv.second.first;
        launches[k] = kv.second.second;
        ++k;
    }
    if ((int)names.size() + 1 > buf_len) return LL_E_CAPACITY;
    memcpy(names_buf, names.c_str(), names.size() + 1);
    return k;
}

// ---- asynchronous submit / collect ------------------------------------------------------------------------------
// ll_submit_scans enqueues the H2D copies of one batch on a copy stream and the whole pipeline behind them on the
// compute stream, then returns; ll_collect blocks for the OLDEST outstanding submission and returns its poses.
// Two submissions may be in flight, so the copies of step k+1 overlap the kernels of step k.  The caller's scan
// buffers must stay valid (and should be pinned) until the matching ll_collect returns.
static int submit_alloc(ll_ctx* c)
{
    if (c->copy_stream) return LL_OK;
    const size_t B = c->B;
    LL_CUDA_CHECK(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    LL_CUDA_CHECK(c, cudaMalloc((void**)&c->d_raw2, sizeof(uint32_t) * B * c->Nmax * 8));
    LL_CUDA_CHECK(c, cudaMalloc((void**)&c->d_hdr2, sizeof(ScanHdr) * B * 2));
    LL_CUDA_CHECK(c, cudaHostAlloc((void**)&c->h_hdr2, sizeof(ScanHdr) * B * 2, cudaHostAllocDefault));
    LL_CUDA_CHECK(c, cudaHostAlloc((void**)&c->h_pose2, sizeof(double) * 2 * B * 14, cudaHostAllocDefault));
    for (int k = 0; k < 2; ++k) {
        LL_CUDA_CHECK(c, cudaEventCreateWithFlags(&c->ev_staged[k], cudaEventDisableTiming));
        LL_CUDA_CHECK(c, cudaEventCreateWithFlags(&c->ev_raw_free[k], cudaEventDisableTiming));
        LL_CUDA_CHECK(c, cudaEventCreateWithFlags(&c->ev_done[k], cudaEventDisableTiming));
    }
    return LL_OK;
}

// the part of a submission after its host-to-device copies have been enqueued on the copy stream
static int submit_launch(ll_ctx* c, int slot, int n_scans, uint32_t* raw, ScanHdr* hdr, ScanHdr* dhdr)
{
    int rc;
    LL_CUDA_CHECK(c, cudaMemcpyAsync(dhdr, hdr, sizeof(ScanHdr) * n_scans, cudaMemcpyHostToDevice, c->copy_stream));
    LL_CUDA_CHECK(c, cudaEventRecord(c->ev_staged[slot], c->copy_stream));
    LL_CUDA_CHECK(c, cudaStreamWaitEvent(c->stream, c->ev_staged[slot], 0));
    k_set_scan_hdr<<<(n_scans + 63) / 64, 64, 0, c->stream>>>(c->d_lane, dhdr, raw, n_scans);
    if (c->prof) ll_prof_harvest(c);
    c->launches = 1;
    if ((rc = ll_launch_features(c, n_scans))) return rc;
    LL_CUDA_CHECK(c, cudaEventRecord(c->ev_raw_free[slot], c->stream));
    if ((rc = ll_launch_odometry(c, n_scans))) return rc;
    if (c->cfg.enable_mapping && (rc = ll_launch_mapping(c, n_scans))) return rc;
    LL_CUDA_CHECK(c, cudaMemcpyAsync(c->h_pose2 + (size_t)slot * c->B * 14, c->d_pose, sizeof(double) * 14 * n_scans, cudaMemcpyDeviceToHost, c->stream));
    LL_CUDA_CHECK(c, cudaMemcpyAsync(c->h_status + (size_t)(1 + slot) * c->B, c->d_status, sizeof(int) * n_scans, cudaMemcpyDeviceToHost, c->stream));
    LL_CUDA_CHECK(c, cudaEventRecord(c->ev_done[slot], c->stream));
    c->sub_n[slot] = n_scans;
    c->sub_count++;
    return LL_OK;
}

int ll_submit_scans(ll_ctx* c, int n_scans, const ll_cloud_view* scans)
{
    if (!c || !scans || n_scans < 1 || n_scans > c->B) return LL_E_INVAL;
    LL_CUDA_CHECK(c, cudaSetDevice(c->dev));
    int rc = submit_alloc(c);
    if (rc) return rc;
    if (c->sub_count >= 2) return LL_E_CAPACITY;  // collect first
    const int slot = (c->sub_head + c->sub_count) & 1;
    uint32_t* raw = slot == 0 ? c->d_raw : c->d_raw2;
    ScanHdr* hdr = reinterpret_cast<ScanHdr*>(c->h_hdr2) + (size_t)slot * c->B;
    ScanHdr* dhdr = reinterpret_cast<ScanHdr*>(c->d_hdr2) + (size_t)slot * c->B;
    for (int i = 0; i < n_scans; ++i) { if ((rc = check_scan_view(c, scans[i]))) return rc; }
    // the slab may still be read by the feature kernels of the submission before last (already collected => done)
    LL_CUDA_CHECK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_raw_free[slot], 0));
    for (int i = 0; i < n_scans; ++i) { if ((rc = stage_one(c, scans[i], raw, i, hdr, c->copy_stream))) return rc; }
    return submit_launch(c, slot, n_scans, raw, hdr, dhdr);
}

int ll_submit_packed(ll_ctx* c, int n_scans, const void* host_base, const int64_t* byte_offsets, const int* n_points, int stride_bytes)
{
    if (!c || !host_base || !byte_offsets || !n_points || n_scans < 1 || n_scans > c->B || !stride_in_place(stride_bytes)) return LL_E_INVAL;
    LL_CUDA_CHECK(c, cudaSetDevice(c->dev));
    int rc = submit_alloc(c);
    if (rc) return rc;
    if (c->sub_count >= 2) return LL_E_CAPACITY;  // collect first
    const int slot = (c->sub_head + c->sub_count) & 1;
    uint32_t* raw = slot == 0 ? c->d_raw : c->d_raw2;
    ScanHdr* hdr = reinterpret_cast<ScanHdr*>(c->h_hdr2) + (size_t)slot * c->B;
    ScanHdr* dhdr = reinterpret_cast<ScanHdr*>(c->d_hdr2) + (size_t)slot * c->B;
    const int64_t lo = byte_offsets[0];
    int64_t end = lo;
    for (int i = 0; i < n_scans; ++i) {
        if (n_points[i] < 1 || byte_offsets[i] < end || (byte_offsets[i] & 3)) return LL_E_INVAL;
        if (n_points[i] > c->Nmax) return LL_E_CAPACITY;
        end = byte_offsets[i] + (int64_t)n_points[i] * stride_bytes;
        const uint64_t off = (uint64_t)(byte_offsets[i] - lo) / 4;
        hdr[i].n_raw = n_points[i]; hdr[i].stride_words = stride_bytes / 4;
        hdr[i].off_lo = (unsigned)off; hdr[i].off_hi = (unsigned)(off >> 32);
    }
    if ((uint64_t)(end - lo) > (uint64_t)c->B * c->Nmax * 32) return LL_E_CAPACITY;   // the staging slab holds B x Nmax x 32 bytes
    LL_CUDA_CHECK(c, cudaStreamWaitEvent(c->copy_stream, c->ev_raw_free[slot], 0));
    LL_CUDA_CHECK(c, cudaMemcpyAsync(raw, static_cast<const char*>(host_base) + lo, (size_t)(end - lo), cudaMemcpyHostToDevice, c->copy_stream));
    return submit_launch(c, slot, n_scans, raw, hdr, dhdr);
}

int ll_collect(ll_ctx* c, double* poses_out)
{
    if (!c) return LL_E_INVAL;
    if (c->sub_count < 1) return LL_E_INVAL;
    LL_CUDA_CHECK(c, cudaSetDevice(c->dev));
    const int slot = c->sub_head;
    LL_CUDA_CHECK(c, cudaEventSynchronize(c->ev_done[slot]));
    if (poses_out) memcpy(poses_out, c->h_pose2 + (size_t)slot * c->B * 14, sizeof(double) * 14 * c->sub_n[slot]);
    c->sub_head ^= 1;
    c->sub_count--;
    const int st = publish_status(c, c->h_status + (size_t)(1 + slot) * c->B, c->sub_n[slot]);
    return st < 0 ? st : c->sub_n[slot];
}

int ll_get_lane_status(ll_ctx* c, int* status, int n)
{
    if (!c || !status || n < 0 || n > c->B) return LL_E_INVAL;
    memcpy(status, c->last_status, sizeof(int) * n);
    return LL_OK;
}

int ll_debug_features(ll_ctx* c, int lane, int counts[5], int* sharp_idx, int* less_sharp_idx, int* flat_idx)
{
    if (!c || lane < 0 || lane >= c->B || !counts) return LL_E_INVAL;
    LL_CUDA_CHECK(c, cudaSetDevice(c->dev));
    LL_CUDA_CHECK(c, cudaMemcpyAsync(c->h_lane, c->d_lane + lane, sizeof(LaneState), cudaMemcpyDeviceToHost, c->stream));
    LL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    const LaneState L = c->h_lane[0];
    counts[0] = L.n_full; counts[1] = L.n_sharp; counts[2] = L.n_less_sharp; counts[3] = L.n_flat; counts[4] = L.n_less_flat;
    cudaStream_t s = c->stream;
    const size_t R = c->R;
    if (sharp_idx && L.n_sharp) LL_CUDA_CHECK(c, cudaMemcpyAsync(sharp_idx, c->d_sharp_idx + lane * R * LL_SHARP_PER_RING, sizeof(int) * L.n_sharp, cudaMemcpyDeviceToHost, s));
    if (less_sharp_idx && L.n_less_sharp) LL_CUDA_CHECK(c, cudaMemcpyAsync(less_sharp_idx, c->d_lsharp_idx + lane * R * LL_LSHARP_PER_RING, sizeof(int) * L.n_less_sharp, cudaMemcpyDeviceToHost, s));
    if (flat_idx && L.n_flat) LL_CUDA_CHECK(c, cudaMemcpyAsync(flat_idx, c->d_flat_idx + lane * R * LL_FLAT_PER_RING, sizeof(int) * L.n_flat, cudaMemcpyDeviceToHost, s));
    LL_CUDA_CHECK(c, cudaStreamSynchronize(s));
    return LL_OK;
}

int ll_last_timings(ll_ctx* c, float ms[4])
{
    if (!c || !ms) return LL_E_INVAL;
    if (c->last_call_graph) { ms[0] = ms[1] = ms[2] = ms[3] = 0.f; return LL_E_INVAL; }   // a replayed graph records no stage events (LL_GRAPH=0 to time stages)
    LL_CUDA_CHECK(c, cudaEventSynchronize(c->ev[3]));
    LL_CUDA_CHECK(c, cudaEventElapsedTime(&ms[0], c->ev[0], c->ev[1]));
    LL_CUDA_CHECK(c, cudaEventElapsedTime(&ms[1], c->ev[1], c->ev[2]));
    LL_CUDA_CHECK(c, cudaEventElapsedTime(&ms[2], c->ev[2], c->ev[3]));
    LL_CUDA_CHECK(c, cudaEventElapsedTime(&ms[3], c->ev[0], c->ev[3]));
    return LL_OK;
}

int ll_extract_features(ll_ctx* c, ll_cloud_view scan, ll_cloud_out* full, ll_cloud_out* sharp, ll_cloud_out* less_sharp, ll_cloud_out* flat,
                        ll_cloud_out* less_flat, int* sharp_idx, int* less_sharp_idx, int* flat_idx, float* curvature, int* ring_begin)
{
    if (!c) return LL_E_INVAL;
    int rc = ll_stage_scans(c, 1, &scan);
    if (rc) return rc;
    c->launches = 0;
    rc = ll_launch_features(c, 1);
    if (rc) return rc;
    rc = fetch_lanes(c, 1);
    if (rc) return rc;
    const LaneState& L = c->h_lane[0];
    if (L.err) return L.err;
    const int cur = L.cur;
    if ((rc = copy_out(c, full, c->d_full, L.n_full))) return rc;
    if ((rc = copy_out(c, sharp, c->d_sharp, L.n_sharp))) return rc;
    if ((rc = copy_out(c, less_sharp, c->d_lsharp[cur], L.n_less_sharp))) return rc;
    if ((rc = copy_out(c, flat, c->d_flat, L.n_flat))) return rc;
    if ((rc = copy_out(c, less_flat, c->d_lflat[cur], L.n_less_flat))) return rc;
    cudaStream_t s = c->stream;
    if (sharp_idx && L.n_sharp) LL_CUDA_CHECK(c, cudaMemcpyAsync(sharp_idx, c->d_sharp_idx, sizeof(int) * L.n_sharp, cudaMemcpyDeviceToHost, s));
    if (less_sharp_idx && L.n_less_sharp) LL_CUDA_CHECK(c, cudaMemcpyAsync(less_sharp_idx, c->d_lsharp_idx, sizeof(int) * L.n_less_sharp, cudaMemcpyDeviceToHost, s));
    if (flat_idx && L.n_flat) LL_CUDA_CHECK(c, cudaMemcpyAsync(flat_idx, c->d_flat_idx, sizeof(int) * L.n_flat, cudaMemcpyDeviceToHost, s));
    if (curvature && L.n_full) LL_CUDA_CHECK(c, cudaMemcpyAsync(curvature, c->d_curv, sizeof(float) * L.n_full, cudaMemcpyDeviceToHost, s));
    if (ring_begin) memcpy(ring_begin, L.ring_begin, sizeof(int) * (c->R + 1));
    LL_CUDA_CHECK(c, cudaStreamSynchronize(s));
    return LL_OK;
}

int ll_fetch_pointcloud2(ll_ctx* c, int which, void* data, int cap_points, int* n_points)
{
    if (!c || which < 0 || which > 4 || (!data && cap_points > 0) || !n_points) return LL_E_INVAL;
    LL_CUDA_CHECK(c, cudaSetDevice(c->dev));
    int rc = fetch_lanes(c, 1);
    if (rc) return rc;
    const LaneState& L = c->h_lane[0];
    const float4* src[5] = {c->d_full, c->d_sharp, c->d_lsharp[L.cur], c->d_flat, c->d_lflat[L.cur]};
    const int cnt[5] = {L.n_full, L.n_sharp, L.n_less_sharp, L.n_flat, L.n_less_flat};
    const int n = cnt[which];
    *n_points = n;
    if (n > cap_points) return LL_E_CAPACITY;
    if (n == 0) return LL_OK;
    if (c->sub_count > 0) return LL_E_INVAL;   // outstanding submissions own the stream order; collect first
    if (!c->d_pc2) LL_CUDA_CHECK(c, cudaMalloc((void**)&c->d_pc2, (size_t)c->Nmax * 32));   // own scratch: staged inputs stay intact
    float4* dst = c->d_pc2;
    k_pack_pointcloud2<<<(n + 255) / 256 < 592 ? (n + 255) / 256 : 592, 256, 0, c->stream>>>(src[which], dst, nullptr, n);
    LL_CUDA_CHECK(c, cudaGetLastError());
    LL_CUDA_CHECK(c, cudaMemcpyAsync(data, dst, (size_t)n * 32, cudaMemcpyDeviceToHost, c->stream));
    LL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return LL_OK;
}

int ll_odometry_step(ll_ctx* c, ll_cloud_view sharp, ll_cloud_view less_sharp, ll_cloud_view flat, ll_cloud_view less_flat, double q_w_curr[4],
                     double t_w_curr[3], double q_last_curr[4], double t_last_curr[3])
{
    if (!c) return LL_E_INVAL;
    LL_CUDA_CHECK(c, cudaSetDevice(c->dev));
    c->launches = 0;
    int rc = fetch_lanes(c, 1);
    if (rc) return rc;
    const int nxt = c->h_lane[0].last_slot ^ 1;  // the slot k_set_feature_counts is about to select
    if ((rc = copy_view_to_device(c, sharp, c->d_sharp, c->R * LL_SHARP_PER_RING))) return rc;
    if ((rc = copy_view_to_device(c, less_sharp, c->d_lsharp[nxt], c->R * LL_LSHARP_PER_RING))) return rc;
    if ((rc = copy_view_to_device(c, flat, c->d_flat, c->R * LL_FLAT_PER_RING))) return rc;
    if ((rc = copy_view_to_device(c, less_flat, c->d_lflat[nxt], c->Nmax))) return rc;
    k_set_feature_counts<<<1, 1, 0, c->stream>>>(c->d_lane, sharp.n, less_sharp.n, flat.n, less_flat.n);
    if ((rc = ll_launch_odometry(c, 1))) return rc;
    if ((rc = fetch_lanes(c, 1))) return rc;
    const LaneState& L = c->h_lane[0];
    if (q_w_curr) memcpy(q_w_curr, L.q_w, sizeof(double) * 4);
    if (t_w_curr) memcpy(t_w_curr, L.t_w, sizeof(double) * 3);
    if (q_last_curr) memcpy(q_last_curr, L.para_q, sizeof(double) * 4);
    if (t_last_curr) memcpy(t_last_curr, L.para_t, sizeof(double) * 3);
    if (L.err) return L.err;
    if (L.now_frame > 1 && L.corner_corr[2] + L.plane_corr[2] < 10) return LL_W_FEW_CORRESPONDENCES;  // LO:814-817
    return LL_OK;
}

int ll_get_last_stats(ll_ctx* c, ll_stats* o)
{
    if (!c || !o) return LL_E_INVAL;
    LL_CUDA_CHECK(c, cudaSetDevice(c->dev));
    const int rc = fetch_lanes(c, 1);
    if (rc) return rc;
    const LaneState& L = c->h_lane[0];
    memset(o, 0, sizeof(*o));
    o->n_full = L.n_full; o->n_sharp = L.n_sharp; o->n_less_sharp = L.n_less_sharp; o->n_flat = L.n_flat; o->n_less_flat = L.n_less_flat;
    for (int k = 0; k < 3; ++k) {
        o->corner_corr[k] = L.corner_corr[k]; o->plane_corr[k] = L.plane_corr[k]; o->plane_selected[k] = L.plane_sel[k];
        o->lm_jacobian_evals[k] = L.jac_evals[k]; o->lm_cost_evals[k] = L.cost_evals[k]; o->lm_termination[k] = L.termination[k];
        o->initial_cost[k] = L.initial_cost[k]; o->final_cost[k] = L.final_cost[k];
    }
    o->map_corner = L.n_map_corner; o->map_surf = L.n_map_surf; o->stack_corner = L.n_stack_corner; o->stack_surf = L.n_stack_surf;
    o->map_corner_corr = L.n_map_corner_corr; o->map_surf_corr = L.n_map_surf_corr;
    for (int k = 0; k < 2; ++k) {
        o->map_jacobian_evals[k] = L.jac_evals[3 + k]; o->map_termination[k] = L.termination[3 + k];
        o->map_initial_cost[k] = L.initial_cost[3 + k]; o->map_final_cost[k] = L.final_cost[3 + k];
    }
    o->frame = L.now_frame;
    o->map_vote_corr = L.n_map_vote; o->map_vote_selected = L.n_map_vote_sel;
    if (getenv("LL_DEBUG_ASSOC")) fprintf(stderr, "assoc dbg (all lanes, cumulative): queued with the 1-NN open %d (kcycles %d, longest %d cycles), queued for the ring window %d (kcycles %d, longest %d cycles)\n", L.dbg[1], L.dbg[3], L.dbg[5] * 16, L.dbg[2], L.dbg[4], L.dbg[6] * 16);
    if (getenv("LL_DEBUG_LM")) fprintf(stderr, "lm dbg (lane, cumulative cycles; LL_LM_TIMING build): linearise %d  reduce %d  step %d  cost %d  reduce %d  accept %d  gradient %d\n", L.dbg[0], L.dbg[1], L.dbg[2], L.dbg[3], L.dbg[4], L.dbg[5], L.dbg[6]);
    o->kernel_launches = c->launches;
    return LL_OK;
}

int ll_debug_assoc(ll_ctx* c, int lane, int* corner, int corner_cap, int* plane, int plane_cap)
{
    if (!c || lane < 0 || lane >= c->B) return LL_E_INVAL;
    LL_CUDA_CHECK(c, cudaSetDevice(c->dev));
    const int nc = c->R * LL_SHARP_PER_RING, np = c->R * LL_FLAT_PER_RING;
    if (corner) LL_CUDA_CHECK(c, cudaMemcpyAsync(corner, c->d_corner_assoc + (size_t)lane * nc * 2, sizeof(int) * 2 * (corner_cap < nc ? corner_cap : nc), cudaMemcpyDeviceToHost, c->stream));
    if (plane) LL_CUDA_CHECK(c, cudaMemcpyAsync(plane, c->d_plane_assoc + (size_t)lane * np * 4, sizeof(int) * 4 * (plane_cap < np ? plane_cap : np), cudaMemcpyDeviceToHost, c->stream));
    LL_CUDA_CHECK(c, cudaStreamSynchronize(c->stream));
    return LL_OK;
}

}  // extern "C"

void ll_prof_harvest(ll_ctx* c)
{
    if (c->prof_name.empty()) return;
    cudaEventSynchronize(c->prof_ev[2 * (c->prof_name.size() - 1) + 1]);
    for (size_t k = 0; k < c->prof_name.size(); ++k) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, c->prof_ev[2 * k], c->prof_ev[2 * k + 1]) == cudaSuccess) {
            auto& a = c->prof_acc[c->prof_name[k]];
            a.first += ms;
            a.second += 1;
        }
    }
    c->prof_name.clear();
}
