// Private context layout of liblightloam_b200 (host + device side).  Not part of the C ABI.
//
// Data layout in HBM (per lane = one independent scan stream; all arrays are [batch][...] slabs so one
// launch covers every lane with blockIdx.y / blockIdx.z = lane):
//   raw        : the caller's point records as uploaded (stride words per point)
//   ring8/rank8: per raw point ring id (int8, -1 = dropped) and rank inside its 256-point tile (the azimuth -atan2(y,x)
//                is recomputed by k_scatter: cheaper than a 4-byte round trip per point)
//   tile_hist  : [tiles][rings] counts -> exclusive offsets (stable counting sort by ring, SR:133-221)
//   full       : ring-sorted float4 x,y,z,intensity (= laserCloud, SR:215-221)
//   curv       : fp32 curvature per full point (SR:225-235)
//   ring lists : per ring picked indices (sharp 12, less-sharp 120, flat 24) + per-ring voxel-DS output
//   compact    : sharp / less_sharp / flat / less_flat clouds in the reference's publish order
//   last[2]    : ping-pong copies of less_sharp / less_flat = laserCloudCornerLast / SurfLast (LO:882-891)
//   index      : polar index (azimuth bin x ring: bucket_start + bucket-sorted float4 with the original index in .w) that
//                replaces kdtreeCornerLast / kdtreeSurfLast (LO:895-896); hashed uniform grids replace the map kd-trees
//   assoc/blocks: correspondence indices and fp64 residual-block records for the LM solve
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/lightloam_b200.h"

#define LL_TILE 256          // points per counting-sort tile
#define LL_MAX_RINGS 64
#define LL_SHARP_PER_RING 12     // 6 sectors x 2   (SR:270-275)
#define LL_LSHARP_PER_RING 120   // 6 sectors x 20  (SR:276-280)
#define LL_FLAT_PER_RING 24      // 6 sectors x 4   (SR:327-331)
#define LL_BLOCK_DOUBLES 12      // residual-block record: type, cp[3], p0[3], p1[3], w, pad

struct LaneState {
    // --- scan registration ---
    const uint32_t* raw;             // this lane's raw point records (staging slab or scan pool)
    int n_raw, stride_words;
    int first_valid, last_valid;     // first / last index surviving the NaN + range filters (SR:109-110)
    int half_idx;                    // index of the point that flips halfPassed (SR:189-192), INT_MAX if none
    float start_ori, end_ori;        // SR:114-126
    float start_dir[2];              // (cos, sin) of the first valid point's azimuth (= -startOri): the halfPassed prefilter of k_classify
    int n_full;
    int ring_begin[LL_MAX_RINGS + 1];
    int n_sharp, n_less_sharp, n_flat, n_less_flat;
    int less_flat_ring_begin[LL_MAX_RINGS + 1];
    // --- odometry ---
    int inited, now_frame;           // LO:33, LO:377
    int cur;                         // ping-pong slot written by the latest extraction / upload (= last_slot ^ 1)
    int last_slot;                   // slot holding laserCloudCornerLast / SurfLast (registered by k_odom_finalize)
    int n_last_corner, n_last_surf;  // sizes of laserCloudCornerLast / laserCloudSurfLast
    int mono_corner, mono_surf;      // 1: the *Last cloud is ring-monotone (int(intensity) non-decreasing, 0..255)
    double para_q[4], para_t[3];     // LO:61-62 (x,y,z,w)
    double q_w[4], t_w[3];           // LO:57-58
    int n_corner_corr, n_plane_corr, n_plane_sel, n_blocks;
    // solver summary per outer iteration (3 odometry + 2 mapping)
    double initial_cost[5], final_cost[5];
    int jac_evals[5], cost_evals[5], termination[5];
    int corner_corr[3], plane_corr[3], plane_sel[3];
    // --- mapping ---
    double map_par[7];               // LM:81 parameters: q_w_curr (xyzw), t_w_curr
    double q_wmap_wodom[4], t_wmap_wodom[3];  // LM:87-88
    double map_odom[7];              // q_wodom_curr, t_wodom_curr of the frame being mapped (LM:90-91)
    int cen[3];                      // laserCloudCenWidth/Height/Depth LM:42-44
    int map_frame;
    int map_shift[3];                // this frame's cube shift: new logical index = old + shift (LM:1596-1779)
    int map_center[3];               // centerCubeI/J/K after shifting
    int n_valid;                     // laserCloudValidNum (<= 125)
    int map_ok;                      // LM:1826 guard: map corner > 10 && map surf > 50
    int n_map_corner, n_map_surf, n_stack_corner, n_stack_surf, n_map_corner_corr, n_map_surf_corr;
    int n_map_vote, n_map_vote_sel;  // plane correspondences the LM:2057-2072 vote saw / selected (last iteration)
    int err;                         // sticky device-side error (LL_E_*)
    int dbg[8];                      // association statistics (plane queries resolved per shell / by the walk)
};

struct KnnGrid {
    int T = 0;            // buckets per lane (power of two)
    int cap = 0;          // points per lane
    float h = 1.f, inv_h = 1.f;
    int* start = nullptr;     // [B][T+1] exclusive bucket offsets
    int* cursor = nullptr;    // [B][T]   counting / scatter cursors
    int* partial = nullptr;   // [B][T / 2048] chunk sums for the two-level scan
    float4* sorted = nullptr; // [B][cap] bucket-ordered points, .w = original index bits
};

struct ll_ctx {
    ll_config cfg;
    int B = 1, R = 64, Nmax = 0, NT = 0, RCAP = 0, SCAP = 0, KCAP = 0;
    int dev = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    std::string last_error;
    int cluster_ok[3] = {-1, -1, -1};   // can the device run the solve kernels (odometry, odometry de-skew, mapping) as clusters? -1 = not asked yet
    int launches = 0;
    int pre_launches = 0;     // launches issued by the staging half of a call, folded into `launches` by the processing half
    float vote_t_min = 0.f;   // smallest fp32 t with expf(-t) < 0.96f on this host's libm (LO:239-242)
    float vote_t95 = 0.f;     // smallest fp32 t with (double)expf(-t) < 0.95 (LM:918-924: float score against a double literal)

    LaneState* d_lane = nullptr;
    LaneState* h_lane = nullptr;   // pinned mirror
    double* h_pose = nullptr;      // pinned [B][14]
    double* d_pose = nullptr;      // [B][14]
    int* h_hdr = nullptr;          // pinned [B][4] {n_raw, stride_words, word offset lo, hi}
    int* d_hdr = nullptr;          // [B][4]
    int* d_status = nullptr;       // [B] per-lane status of the last step (LaneState::err at k_odom_finalize / k_map_end)
    int* h_status = nullptr;       // pinned [3][B]: slot 0 = synchronous calls, 1..2 = submit slots
    int* last_status = nullptr;    // host [B]: what ll_get_lane_status reports
    bool feat_attr_set = false;    // dynamic shared-memory limits of the per-ring kernels set on this context's device
    int* d_wide_list = nullptr;    // [B * R] rings longer than 3083 points of the current step (lane * R + ring)
    int* d_wide_n = nullptr;       // [1]
    float4* d_pc2 = nullptr;       // scratch of ll_fetch_pointcloud2 (allocated on first use)

    uint32_t* d_raw = nullptr;     // [B][Nmax * 8] words (stride <= 32 B)
    int8_t* d_ring8 = nullptr;     // [B][Nmax]
    uint8_t* d_rank8 = nullptr;    // [B][Nmax]
    int* d_tile_hist = nullptr;    // [B][NT][R]
    float4* d_full = nullptr;      // [B][Nmax]
    float* d_curv = nullptr;       // [B][Nmax]
    int8_t* d_label = nullptr;     // [B][Nmax] cloudLabel (SR:40)
    unsigned* d_brk = nullptr;     // [B][R][(RCAP+31)/32] consecutive-gap break bits (SR:290-293)
    uint16_t* d_sorted16 = nullptr;// [B][Nmax] per-sector sorted local index | curvature class bits
    float4* d_lf_tmp = nullptr;    // [B][Nmax] per-ring voxel-DS output parked at the ring's own offset
    int* d_ring_lists = nullptr;   // [B][R][12 + 120 + 24]
    int* d_ring_counts = nullptr;  // [B][R][4]  sharp, less_sharp, flat, less_flat
    float4* d_sharp = nullptr;     // [B][R*12]
    float4* d_flat = nullptr;      // [B][R*24]
    int* d_sharp_idx = nullptr;    // [B][R*12]
    int* d_lsharp_idx = nullptr;   // [B][R*120]
    int* d_flat_idx = nullptr;     // [B][R*24]
    float4* d_lsharp[2] = {nullptr, nullptr};  // [B][R*120]
    float4* d_lflat[2] = {nullptr, nullptr};   // [B][Nmax]

    KnnGrid a_corner, a_surf;      // polar (azimuth bin x ring) index over last less-sharp / less-flat (T = rings * az_bins)
    unsigned* d_ebound[2] = {nullptr, nullptr};  // [B][R][2] per-ring elevation band of the indexed cloud (order-preserving encodings)
    float* d_bands[2] = {nullptr, nullptr};      // [B][2 * LL_MAX_RINGS + 4] the same decoded (k_index_partial)
    int az_bins_corner = 64, az_bins_surf = 256;
    int* d_corner_assoc = nullptr; // [B][R*12][2]
    int* d_plane_assoc = nullptr;  // [B][R*24][4]
    float4* d_vote_src = nullptr;  // [B][R*24] compacted plane matches for the graph vote (current point, w = feature index)
    float4* d_vote_tgt = nullptr;  // [B][R*24] ... and their closest points
    int4* d_assoc_queue = nullptr; // [B * R * 36] queries handed to the warp pass of the association
    int* d_assoc_queue_n = nullptr;// [16] per outer iteration: long entries queued [0..2], pop cursors [4..6], short entries queued [8..10]
    int assoc_queue_cap = 0;
    float4* d_qa = nullptr;        // [B][R*36] transformed queries of the outer iteration, grouped by azimuth slab (k_odom_queries)
    float4* d_qb = nullptr;        // [B][R*36] their polar coordinates
    int* d_qstart = nullptr;       // [B][qstart_stride] slab offsets
    int qstart_stride = 260;
    bool slab_attr_set = false, vp_attr_set = false;
    // split odometry solve (few lanes, many SMs): `parts` CTAs per lane all-reduce the 28 doubles through a local mailbox (LmComm)
    void* d_odom_comm = nullptr;               // mailbox doubles followed by the sequence flags
    size_t odom_comm_mbox_bytes = 0;
    unsigned long long* d_odom_seq[2] = {nullptr, nullptr};   // [B] collective counters, alternated per solve launch
    int odom_comm_flip = 0;
    double* d_blocks = nullptr;    // [B][nblk_cap][12]
    int nblk_cap = 0;

    // asynchronous submit / collect: a second staging slab + a copy stream so the H2D of step k+1 overlaps the
    // kernels of step k (ll_submit_scans / ll_collect)
    cudaStream_t copy_stream = nullptr;
    uint32_t* d_raw2 = nullptr;        // second staging slab [B][Nmax * 8]
    int* h_hdr2 = nullptr;             // pinned headers per slot [2][B][4]
    int* d_hdr2 = nullptr;             // [2][B][4]
    double* h_pose2 = nullptr;         // pinned poses per slot [2][B][14]
    cudaEvent_t ev_staged[2] = {nullptr, nullptr}, ev_raw_free[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    int sub_n[2] = {0, 0};             // scans of the submission occupying each slot
    int sub_head = 0, sub_count = 0;   // ring of outstanding submissions (at most 2)

    // scan pool: scans kept resident in HBM and referenced by id (ll_pool_upload / ll_process_pool)
    uint32_t* d_pool = nullptr;    // [pool_cap][Nmax * 4] words (float4 records)
    int* d_pool_n = nullptr;       // [pool_cap] points per pooled scan
    int* h_ids = nullptr;          // pinned [B]
    int* d_ids = nullptr;          // [B]
    int pool_cap = 0, pool_n = 0;

    // CUDA graphs of the fused pipeline for the latency path (few lanes): key = lanes * 2 + parity of the solve's mailbox counters
    std::map<int, cudaGraphExec_t> graphs;
    std::map<int, int> graph_launches;   // kernels in each captured graph
    int graph_parity_step = 0;     // how one step moves odom_comm_flip (1 when the solves are split, else 0)
    int eager_calls = 0;           // the first call of a context runs eagerly (function attributes, lazy allocations)
    bool last_call_graph = false;  // ll_last_timings has no events to read after a graph launch

    // optional per-kernel device timing (ll_profile_enable): event pairs around every launch
    bool prof = false;
    std::vector<cudaEvent_t> prof_ev;
    std::vector<const char*> prof_name;      // name per pair in flight
    std::map<std::string, std::pair<double, int>> prof_acc;  // name -> (total ms, launches)

    // mapping (allocated when enable_mapping)
    struct MapState* map = nullptr;
};

// error plumbing ---------------------------------------------------------------------------------------
#define LL_CUDA_CHECK(ctx, expr)                                                                      \
    do {                                                                                              \
        cudaError_t e__ = (expr);                                                                     \
        if (e__ != cudaSuccess) {                                                                     \
            (ctx)->last_error = std::string(#expr) + ": " + cudaGetErrorString(e__);                  \
            return LL_E_CUDA;                                                                         \
        }                                                                                             \
    } while (0)

// RAII event pair around one kernel launch when profiling is on
struct LLProf {
    ll_ctx* c;
    bool on;
    LLProf(ll_ctx* ctx, const char* name) : c(ctx), on(ctx->prof)
    {
        if (!on) return;
        const size_t k = c->prof_name.size();
        if (c->prof_ev.size() < 2 * (k + 1)) {
            cudaEvent_t a, b;
            cudaEventCreate(&a);
            cudaEventCreate(&b);
            c->prof_ev.push_back(a);
            c->prof_ev.push_back(b);
        }
        c->prof_name.push_back(name);
        cudaEventRecord(c->prof_ev[2 * k], c->stream);
    }
    ~LLProf()
    {
        if (on) cudaEventRecord(c->prof_ev[2 * (c->prof_name.size() - 1) + 1], c->stream);
    }
};
void ll_prof_harvest(ll_ctx* c);  // folds the in-flight event pairs into prof_acc (synchronises)

// kernel-group entry points (defined in the .cu files) ---------------------------------------------------
int ll_launch_features(ll_ctx* c, int n_lanes);                 // SR:100-377 on lanes [0, n_lanes)
int ll_launch_odometry(ll_ctx* c, int n_lanes);                 // LO:425-896
int ll_map_alloc(ll_ctx* c);
void ll_map_free(ll_ctx* c);
void ll_map_clear(ll_ctx* c);                                   // forget the map contents, keep the allocations
int ll_launch_mapping(ll_ctx* c, int n_lanes);                  // LM:1581-2168
// host-side state a captured mapping frame depends on (which of the two cube-map buffers is current): part of the CUDA graph
// key; -1 = this context's mapping cannot be replayed from a graph (slab sharding over several GPUs, mailbox split)
int ll_map_graph_state(const ll_ctx* c);
void ll_map_graph_set_state(ll_ctx* c, int state);
