// ORACLE — TEST INFRASTRUCTURE ONLY.  CPU restatement of the Light-LOAM per-scan hot path.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build,
// link or call anything under oracle/.  The product path (light-loam_b200/) never includes these
// files and fails loudly when its CUDA library is missing.
//
// PARITY UNPINNED: the reference (BrenYi/Light-LOAM @ 78aa294) ships no tests, golden vectors or
// fixtures, and its sources cannot be compiled here (ROS, PCL/FLANN, Ceres, Eigen are absent from
// the image and un-vendored), so this restatement cannot be checked against reference outputs.
// It follows the reference sources line by line (citations below) and restates the published
// algorithms of PCL 1.10 VoxelGrid / KdTreeFLANN, Ceres 2.x TrustRegionMinimizer + HuberLoss +
// EigenQuaternionManifold + DENSE_QR + Jet autodiff, and Eigen 3.3 quaternion / slerp / 3x3
// eigen / 5x3 least squares.  Independent cross-checks live in tests/ (brute-force k-NN, SciPy
// cKDTree, NumPy finite-difference Jacobians and LM step).
#pragma once
#include <cstdint>
#include <vector>

namespace orc {

// pcl::PointXYZI as used by the reference (common.h:7). 16 B here (PCL pads to 32 B).
struct P4 { float x, y, z, i; };

// ROS params + compile-time constants of the reference (SURVEY.md §5 "Config / flags").
struct Config {
    int   scan_line     = 64;      // scanRegistration.cpp:435
    float minimum_range = 5.0f;    // scanRegistration.cpp:438 (double param, passed as float thres SR:110)
    float lower_bound   = -24.9f;  // scanRegistration.cpp:439
    float up_bound      = 2.0f;    // scanRegistration.cpp:440
    float line_res      = 0.4f;    // laserMapping.cpp:2363
    float plane_res     = 0.8f;    // laserMapping.cpp:2364
    int   skip_frame    = 1;       // laserOdometry.cpp:350
    int   voxel_stable  = 0;       // 0: std::sort like PCL (reference-faithful, within-voxel order
                                   //    implementation-defined); 1: stable input order (what the GPU does)
    int   graph_from_frame = 5;    // laserOdometry.cpp:781,794: vote when now_frame > 5
    int   map_graph_vote = 0;      // laserMapping.cpp:2057-2072 (commented out in the reference): 0 = off as shipped; N > 0 = the
                                   // block enabled from mapping frame N - 1 on (the reference text reads "now_frame > 20" = 22)
    int   distortion = 0;          // laserOdometry.cpp:23 DISTORTION (reference build: 0)
    int   vote_mode = 0;           // 0: vote_simple (live); 1: vote_partial (laserMapping.cpp:261-834, dead in the reference)
};

struct Features {
    std::vector<P4>    full;             // ring-sorted laserCloud (SR:215-221)
    std::vector<int>   ring_begin;       // scan_line+1 offsets into full
    std::vector<float> curvature;        // SR:231 (0 outside [5, n-5))
    std::vector<int>   label;            // SR:234, 272, 278, 324
    std::vector<int>   sharp_idx, less_sharp_idx, flat_idx;  // indices into full, in push_back order
    std::vector<P4>    sharp, less_sharp, flat, less_flat;
    std::vector<int>   less_flat_ring_count;                 // DS output points per ring
    long               sort_ties = 0;    // tie audit: equal adjacent curvatures inside a sorted sector
};

}  // namespace orc
