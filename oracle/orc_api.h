// ORACLE (test infrastructure) — internal C++ API of the CPU restatement.  See orc_types.h header.
#pragma once
#include "orc_types.h"

#include <array>
#include <memory>

namespace orc {

// ---- scanRegistration.cpp:87-428 ------------------------------------------------------------------
int extract_features(const float* pts, int n_in, int stride_floats, const Config& cfg, Features& out);

// ---- pcl::VoxelGrid<PointXYZI> (SR:370-374, LM:1815-1821, LM:2160-2166) ---------------------------
void voxel_grid(const std::vector<P4>& in, float leaf, bool stable, std::vector<P4>& out);

// ---- pcl::KdTreeFLANN<PointXYZI>: exact sorted k-NN on xyz, fp32 L2_Simple (LO:494,656; LM:1882,1948)
class KdTree {
public:
    void build(const std::vector<P4>& pts);                         // setInputCloud
    int  knn(const float q[3], int k, int* idx, float* d2) const;   // nearestKSearch; returns #found
    bool empty() const { return n_ == 0; }
private:
    struct Node { int left, right, begin, end; float lo[3], hi[3]; };
    int build_rec(int begin, int end);
    void search(int node, const float q[3], int k, int* idx, float* d2, int& found) const;
    std::vector<Node> nodes_;
    std::vector<int> order_;
    std::vector<float> xyz_;  // reordered copy, 3 floats per point
    int n_ = 0;
};

// ---- lidarFactor.hpp live functors + ceres::Solve (LO:475-482,819-825; LM:1865-1872,2079-2087) ----
enum BlockType { EDGE = 0, PLANE_MODIFY = 1, PLANE_NORM = 2 };
struct ResidualBlock {
    int type;
    double cp[3];      // curr_point
    double a[3];       // EDGE: last_point_a | PLANE_MODIFY: last_point_j | PLANE_NORM: plane_unit_norm
    double b[3];       // EDGE: last_point_b | PLANE_MODIFY: ljm_norm (constructor, LF:210-211)
    double s;          // interpolation ratio (1.0: DISTORTION 0)
    double w;          // PLANE_MODIFY: weight | PLANE_NORM: negative_OA_dot_norm
};
ResidualBlock make_edge(const double cp[3], const double a[3], const double b[3], double s);
ResidualBlock make_plane_modify(const double cp[3], const double j[3], const double l[3], const double m[3], double s, double weight);
ResidualBlock make_plane_norm(const double cp[3], const double n[3], double d);

struct IterRecord { double cost, cost_change, gradient_max_norm, step_norm, relative_decrease, radius; int valid, successful; };
struct SolveSummary {
    double initial_cost = 0, final_cost = 0;
    int num_iterations = 0;          // size of iterations list incl. iteration 0
    int num_jacobian_evals = 0, num_cost_evals = 0;
    int termination = 0;             // 0 no-convergence(max iters) 1 gradient 2 parameter 3 function 4 radius 5 failure
    std::vector<IterRecord> iterations;
};
// Evaluates cost (+ optionally residuals, gradient (6), dense row-major n_rows x 6 Jacobian) at x = (q xyzw, t).
int  evaluate(const std::vector<ResidualBlock>& blocks, const double x[7], double* cost, std::vector<double>* residuals,
              double* gradient6, std::vector<double>* jacobian, bool use_autodiff);
void manifold_plus(const double x[7], const double delta[6], double out[7]);
// ceres::Solve with HuberLoss(0.1), EigenQuaternionManifold, DENSE_QR, max_num_iterations = 4.
void solve(const std::vector<ResidualBlock>& blocks, double q[4], double t[3], SolveSummary* summary,
           int max_num_iterations = 4, bool use_autodiff = true);

// ---- laserOdometry.cpp ------------------------------------------------------------------------------
struct VertexVote { int index; float score; };   // common.h:40-43
struct CorreMatch { int index; P4 src, tgt; float score, s; };  // common.h:20-31
// graph_based_correspondence_vote_simple (LO:165-342), plane case (corner_case=false) or corner case.
void graph_vote_simple(const std::vector<CorreMatch>& correspondences, bool corner_case, std::vector<VertexVote>& selected,
                       std::vector<float>* votes_out /* per correspondence, optional */);

// graph_based_correspondence_vote_partial (LM:321-834, dead in the reference): optional scoring mode, cfg.vote_mode = 1
void graph_vote_partial(const std::vector<CorreMatch>& correspondences, bool corner_case, std::vector<VertexVote>& selected_idx);
// the copy laserMapping.cpp carries (LM:836-1027): 20 regions, threshold 0.95, corner_case branch selects votes < 0.75 m
void graph_vote_simple_mapping(const std::vector<CorreMatch>& correspondences, bool corner_case, std::vector<VertexVote>& selected_idx);

struct OdomIterStats { int corner_corr, plane_corr, plane_selected; SolveSummary solve; };
struct Odometry {
    Config cfg;
    bool systemInited = false;
    int now_frame = 0;
    double para_q[4] = {0, 0, 0, 1};   // LO:61
    double para_t[3] = {0, 0, 0};      // LO:62
    double q_w_curr[4] = {0, 0, 0, 1}; // LO:57 (x,y,z,w)
    double t_w_curr[3] = {0, 0, 0};    // LO:58
    std::vector<P4> cornerLast, surfLast;
    KdTree kdCorner, kdSurf;
    std::vector<OdomIterStats> last_stats;       // one per opti_counter of the last step
    // association dump of the last outer iteration of the last step (for parity debugging)
    std::vector<std::array<int, 3>> last_corner_assoc;  // {query i, closest, minPointInd2}
    std::vector<std::array<int, 4>> last_plane_assoc;   // {query i, closest, minPointInd2, minPointInd3}
    // LO:384-929 for one synchronized set of feature clouds.
    void step(const std::vector<P4>& sharp, const std::vector<P4>& less_sharp, const std::vector<P4>& flat,
              const std::vector<P4>& less_flat);
};

// ---- laserMapping.cpp --------------------------------------------------------------------------------
struct Mapping {
    static const int W = 21, H = 21, D = 11, NUM = W * H * D;  // LM:45-50
    Config cfg;
    int cenW = 10, cenH = 10, cenD = 5;                          // LM:42-44
    std::vector<std::vector<P4>> cornerArray, surfArray;         // LM:73-74
    double parameters[7] = {0, 0, 0, 1, 0, 0, 0};                // LM:81 q_w_curr (xyzw), t_w_curr
    double q_wmap_wodom[4] = {0, 0, 0, 1}, t_wmap_wodom[3] = {0, 0, 0};  // LM:87-88
    int frameCount = 0;
    int last_corner_num = 0, last_surf_num = 0, last_map_corner = 0, last_map_surf = 0;
    int last_stack_corner = 0, last_stack_surf = 0;
    int last_vote_selected = 0;                                  // plane correspondences the LM:2057-2072 vote selected (last iteration)
    std::vector<SolveSummary> last_solves;
    Mapping() : cornerArray(NUM), surfArray(NUM) {}
    // LM:1581-2168 for one (corner_last, surf_last, odom pose) triple; returns 0 or 1 (map too small, LM:2097-2100)
    int step(const std::vector<P4>& cornerLast, const std::vector<P4>& surfLast, const double q_wodom_curr[4],
             const double t_wodom_curr[3]);
    // map preload for the config-3 style benchmark: insert map-frame points into their cubes
    void insert_map_points(const std::vector<P4>& corner, const std::vector<P4>& surf);
};

// small dense helpers shared by mapping (Eigen restatements)
void sym_eig3(const double A[9], double evals[3], double evecs[9]);  // ascending, columns = eigenvectors
bool plane_fit5(const double pts[15], double n[3]);                  // colPivHouseholderQr solve A n = -1

}  // namespace orc
