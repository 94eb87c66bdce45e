/* lightloam_b200 — C ABI of the B200-native Light-LOAM per-scan hot path.
 *
 * The reference (BrenYi/Light-LOAM) has no plugin / FFI interface: the path is inlined in three ROS
 * node bodies.  Each entry point below replaces one of those in-process bodies, cited by
 * reference file:line (scanRegistration.cpp = SR, laserOdometry.cpp = LO, laserMapping.cpp = LM):
 *
 *   ll_extract_features   SR:100-377   everything between PointCloud2 decode and the 5 publishes
 *   ll_odometry_step      LO:425-896   one pass of the odometry loop body incl. warm start + kd rebuild
 *   ll_mapping_step       LM:1581-2168 one pass of process() between message decode and publish
 *   ll_process_scans      the three chained on the device for `batch` independent scan streams
 *
 * Plain C99: opaque context, POD config, raw pointers + sizes, integer error codes, no exceptions.
 * All `float*` clouds are packed x,y,z,intensity fp32 (16 B per point) unless a stride is given.
 * A context is not thread-safe (the reference drives each node from one thread: SR:475, LO:380, LM:2397);
 * different contexts are independent.  Every call is synchronous w.r.t. the context's CUDA stream
 * unless stated otherwise.  There is no CPU fallback: without a CUDA device ll_create fails.
 */
#ifndef LIGHTLOAM_B200_H
#define LIGHTLOAM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LL_OK 0
#define LL_E_INVAL (-1)      /* bad argument */
#define LL_E_CAPACITY (-2)   /* input larger than the capacities given in ll_config */
#define LL_E_CUDA (-3)       /* CUDA runtime error (message via ll_last_error) */
#define LL_E_NCCL (-4)
#define LL_E_EMPTY (-5)      /* no valid point in the scan (the reference would index points[0], SR:114) */
#define LL_W_FEW_CORRESPONDENCES 1 /* warning, mirrors LO:814-817 / LM:2097-2100; results are still valid */

typedef struct ll_ctx ll_ctx;

/* ROS params + compile-time constants of the reference, one POD. */
typedef struct ll_config {
    int   scan_line;         /* SR:435  16 / 32 / 64 */
    float minimum_range;     /* SR:438 */
    float lower_bound;       /* SR:439  deg, 64-line formula */
    float up_bound;          /* SR:440 */
    float line_res;          /* LM:2363 mapping_line_resolution */
    float plane_res;         /* LM:2364 mapping_plane_resolution */
    int   skip_frame;        /* LO:350 (publication decimation only; kept for the nodes) */
    int   graph_from_frame;  /* LO:781,794: planes are graph-voted when now_frame > this (5) */
    int   device;            /* CUDA device ordinal */
    int   batch;             /* number of independent scan streams ("lanes") processed per call */
    int   max_points;        /* capacity: points per raw scan */
    int   max_ring_points;   /* capacity: points in one ring after filtering */
    int   map_capacity;      /* capacity: points per map cloud (corner / surf) in the 5x5x3 local map */
    int   enable_mapping;    /* 0: ll_process_scans stops after odometry */
    int   map_graph_vote;    /* 1: graph_based_correspondence_vote_simple on the scan-to-map plane correspondences (the call the
                              * reference keeps commented out at LM:2057-2072; BASELINE.json configs[2] "graph matching on") */
    int   distortion;        /* LO:23 DISTORTION: 1 = motion de-skew (per-point s = relTime, slerp, TransformToEnd); reference build: 0 */
    int   vote_mode;         /* 0: graph_based_correspondence_vote_simple (LO:165-342, the live code); 1: paper-style
                              * graph_based_correspondence_vote_partial (LM:261-834, dead in the reference; beyond-reference) */
    int   reserved[3];
} ll_config;

typedef struct ll_cloud_view {   /* caller-owned HOST memory */
    const float* data;
    int n;
    int stride_bytes;            /* >= 12; x,y,z (fp32) at offsets 0,4,8 (PointCloud2 point_step).  Raw scans: any value (records
                                  * wider than 32 bytes or not a multiple of 4, e.g. 22-byte XYZIRT or 48-byte Ouster points, are
                                  * gathered to packed xyz by a strided copy).  Feature clouds: multiple of 4.  Where the
                                  * intensity matters (feature clouds of ll_odometry_step / ll_mapping_step): byte 12 when
                                  * stride_bytes < 32 (packed float4), byte 16 when stride_bytes >= 32 (pcl::PointXYZI as
                                  * PointCloud2 carries it) */
} ll_cloud_view;

typedef struct ll_cloud_out {    /* caller-owned HOST memory, float4 per point */
    float* xyzi;
    int n;                       /* out: points written */
    int cap;                     /* in: capacity in points */
} ll_cloud_out;

typedef struct ll_stats {
    int n_full, n_sharp, n_less_sharp, n_flat, n_less_flat;     /* last ll_extract_features / lane 0 */
    int corner_corr[3], plane_corr[3], plane_selected[3];       /* per outer iteration of the last odometry step */
    int lm_jacobian_evals[3], lm_cost_evals[3], lm_termination[3];
    double initial_cost[3], final_cost[3];
    int map_corner, map_surf, stack_corner, stack_surf, map_corner_corr, map_surf_corr;
    int map_jacobian_evals[2], map_termination[2];
    double map_initial_cost[2], map_final_cost[2];
    int frame;                                                   /* now_frame of lane 0 */
    int kernel_launches;                                         /* launches issued by the last call */
    int map_vote_corr, map_vote_selected;                        /* scan-to-map graph vote (map_graph_vote): planes seen / selected, last iteration */
} ll_stats;

void ll_default_config(ll_config* cfg, int scan_line);   /* launch-file values for 16 / 32 / 64 lines */
int  ll_create(const ll_config* cfg, ll_ctx** out);
void ll_destroy(ll_ctx* ctx);
const char* ll_strerror(int code);
const char* ll_last_error(const ll_ctx* ctx);            /* detail of the last LL_E_CUDA */
int  ll_get_last_stats(ll_ctx* ctx, ll_stats* out);
int  ll_reset(ll_ctx* ctx);                              /* forget all streams' state (poses, last clouds, map) */

/* SR:100-377.  Outputs may be NULL.  sharp_idx / less_sharp_idx / flat_idx are indices into `full`
 * (capacities: 12 / 120 / 24 per ring).  Runs on lane 0. */
int ll_extract_features(ll_ctx* ctx, ll_cloud_view scan, ll_cloud_out* full, ll_cloud_out* sharp, ll_cloud_out* less_sharp,
                        ll_cloud_out* flat, ll_cloud_out* less_flat, int* sharp_idx, int* less_sharp_idx, int* flat_idx,
                        float* curvature /* n_full floats or NULL */, int* ring_begin /* scan_line+1 or NULL */);

/* Wire format of the published clouds: packs cloud `which` of lane 0's last extraction (0 = full cloud /velodyne_cloud_2,
 * 1 = /laser_cloud_sharp, 2 = /laser_cloud_less_sharp, 3 = /laser_cloud_flat, 4 = /laser_cloud_less_flat; SR:382-410) on
 * the device into pcl::PointXYZI records exactly as pcl::toROSMsg lays them out in sensor_msgs/PointCloud2::data
 * (point_step 32: x@0, y@4, z@8, intensity@16, little-endian fp32, padding zeroed) and copies them to `data`
 * (cap_points * 32 bytes).  *n_points = points in the cloud (also when LL_E_CAPACITY is returned).  The input
 * direction needs no call: ll_cloud_view::stride_bytes = point_step reads PointCloud2 / KITTI .bin records in place. */
int ll_fetch_pointcloud2(ll_ctx* ctx, int which, void* data, int cap_points, int* n_points);

/* LO:425-896 on lane 0 for one synchronized set of the four feature clouds (float4 xyzi, intensity =
 * ring + 0.1*relTime).  Quaternions are x,y,z,w.  The first call only initialises (LO:427-431). */
int ll_odometry_step(ll_ctx* ctx, ll_cloud_view sharp, ll_cloud_view less_sharp, ll_cloud_view flat, ll_cloud_view less_flat,
                     double q_w_curr[4], double t_w_curr[3], double q_last_curr[4], double t_last_curr[3]);

/* LM:1581-2168 on lane 0. */
int ll_mapping_step(ll_ctx* ctx, ll_cloud_view corner_last, ll_cloud_view surf_last, const double q_wodom_curr[4],
                    const double t_wodom_curr[3], double q_w_curr[4], double t_w_curr[3]);
/* Pre-loads map-frame points into lane 0's cube map (config-3 style benchmarks; no reference equivalent). */
int ll_map_insert(ll_ctx* ctx, ll_cloud_view corner, ll_cloud_view surf);

/* Multi-GPU scan-to-map (BASELINE.json configs[4]: map in spatial slabs, per-GPU JtJ, all-reduce of the 6x6 normal
 * equations; no reference equivalent — the reference is single-process).  One context per GPU / process.  Every rank
 * calls ll_mapping_step with the SAME clouds and odometry pose; rank r only associates the stack points whose
 * map-frame x lies in its slab [x_lo, x_hi) (LM:1877-2055 for those points), and the LM kernel all-reduces the 28
 * doubles (21 JtJ + 6 Jtr + cost) of every linearisation over peer memory inside the kernel, so all ranks take the
 * identical step and return the identical pose.  ll_comm_export / ll_comm_attach(kind 0) connect processes through
 * CUDA IPC handles (exchange them with any host-side all-gather); kind 1 connects contexts of one process through
 * ll_comm_local_ptr.  Attach right after ll_create (all ranks at the same point of identical call sequences). */
#define LL_COMM_HANDLE_BYTES 64
int   ll_comm_export(ll_ctx* ctx, void* handle_out /* LL_COMM_HANDLE_BYTES */);
void* ll_comm_local_ptr(ll_ctx* ctx);
int   ll_comm_attach(ll_ctx* ctx, int rank, int world, const void* peers /* world entries */, int kind);
int   ll_comm_detach(ll_ctx* ctx);
int   ll_map_set_slab(ll_ctx* ctx, double x_lo, double x_hi);

/* Fused device pipeline: scan i of the call feeds lane i (n_scans <= batch).  Per lane the call is
 * SR:100-377 -> LO:425-896 (-> LM:1581-2168 when enable_mapping), with features handed over in HBM.
 * poses_out: n_scans x 14 doubles = odometry q_w_curr[4], t_w_curr[3], mapped q_w_curr[4], t_w_curr[3]
 * (mapped = odometry when mapping is off). */
int ll_process_scans(ll_ctx* ctx, int n_scans, const ll_cloud_view* scans, double* poses_out);
/* The same split in two so a caller can keep inputs resident in HBM: stage = H2D only, run = kernels
 * (+ D2H of the poses when poses_out != NULL). */
int ll_stage_scans(ll_ctx* ctx, int n_scans, const ll_cloud_view* scans);
int ll_process_staged(ll_ctx* ctx, int n_scans, double* poses_out);

/* Asynchronous form of ll_process_scans: ll_submit_scans enqueues the H2D copies (copy stream) and the whole pipeline
 * behind them (compute stream) and returns; ll_collect blocks for the OLDEST outstanding submission, writes its poses
 * (n x 14 doubles) and returns the number of scans it held.  At most two submissions may be in flight, which lets
 * the copies of step k+1 overlap the kernels of step k.  Scan buffers must stay valid until the matching collect. */
int ll_submit_scans(ll_ctx* ctx, int n_scans, const ll_cloud_view* scans);
int ll_collect(ll_ctx* ctx, double* poses_out);
/* ll_submit_scans for scans that lie in ONE caller-owned (pinned) host buffer: scan i = n_points[i] records of stride_bytes
 * (12 = packed xyz, what scanRegistration consumes, SR:105-110; 16 = KITTI x,y,z,i; <= 32, multiple of 4) starting
 * byte_offsets[i] bytes (multiple of 4, ascending, non-overlapping) after host_base.  The whole batch crosses the bus as one
 * host-to-device copy of [byte_offsets[0], end of the last scan). */
int ll_submit_packed(ll_ctx* ctx, int n_scans, const void* host_base, const int64_t* byte_offsets, const int* n_points, int stride_bytes);
/* Per-lane status of the last completed call / collect (n entries): 0, or the LL_E_* code that made the lane skip the scan
 * (LL_E_EMPTY, LL_E_CAPACITY, LL_E_NCCL).  A lane with a non-zero status did not advance: pose, *Last clouds and map unchanged.
 * The batch calls (ll_process_*, ll_collect) also return the first non-zero lane status as their (negative) return code
 * after writing every pose. */
int ll_get_lane_status(ll_ctx* ctx, int* status, int n);

/* Scan pool: upload many scans once (float4 records in HBM), then feed lane i from pooled scan scan_ids[i].
 * Same per-lane semantics as ll_process_scans without the per-call H2D of the points. */
int ll_pool_upload(ll_ctx* ctx, int n_scans, const ll_cloud_view* scans);
int ll_process_pool(ll_ctx* ctx, int n_lanes, const int* scan_ids, double* poses_out);

/* Per-kernel device timing: CUDA event pairs around every launch while enabled. ll_profile_read returns the
 * number of kernel names written ('\n'-separated) with their accumulated milliseconds and launch counts. */
int ll_profile_enable(ll_ctx* ctx, int on);
int ll_profile_read(ll_ctx* ctx, char* names_buf, int buf_len, double* total_ms, int* launches, int cap);

/* Device-side timing of the last ll_process_staged / ll_process_scans: milliseconds between CUDA
 * events recorded on the context stream around the feature, odometry and mapping kernel groups. */
int ll_last_timings(ll_ctx* ctx, float ms[4] /* features, odometry, mapping, total */);
/* Parity / debugging: association indices of the last outer iteration of lane `lane`.
 * corner: n_sharp x {closest, minPointInd2} ; plane: n_flat x {closest, minPointInd2, minPointInd3, weight*1000}; -1 = none */
int ll_debug_assoc(ll_ctx* ctx, int lane, int* corner, int corner_cap, int* plane, int plane_cap);
/* Parity: feature indices (into the lane's ring-sorted cloud) of the last extraction of lane `lane`.
 * counts = {n_full, n_sharp, n_less_sharp, n_flat, n_less_flat}; index arrays sized scan_line * 12 / 120 / 24 (or NULL). */
int ll_debug_features(ll_ctx* ctx, int lane, int counts[5], int* sharp_idx, int* less_sharp_idx, int* flat_idx);
/* Parity of the map filter's device-wide primitives (csrc/ll_sort.cuh) on their own: sorts n (key, value) pairs by the low
 * `key_bits` bits of the key, stable, in place on the host arrays; if scan_io is not NULL it is replaced by its exclusive
 * prefix sum (n_scan ints).  The device arrays are sized for `capacity` >= n elements, as in the filter, so the kernels'
 * device-side length handling is exercised too. */
int ll_debug_sort_scan(ll_ctx* ctx, unsigned long long* keys_io, int* vals_io, int n, int key_bits, int capacity, int* scan_io, int n_scan);
/* Raw CUDA stream of the context (cudaStream_t as void*), so a host can order its own work after ours. */
void* ll_cuda_stream(ll_ctx* ctx);
/* Kernels of this library launched by the last call (no synchronisation; the same number ll_stats::kernel_launches reports). */
int ll_launch_count(const ll_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* LIGHTLOAM_B200_H */
