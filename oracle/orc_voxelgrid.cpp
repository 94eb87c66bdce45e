// ORACLE (test infrastructure) — pcl::VoxelGrid<pcl::PointXYZI>::applyFilter restated.
//
// PCL is an un-vendored dependency of the reference (CMakeLists.txt:22 "PCL 1.10"); call sites:
// scanRegistration.cpp:370-374 (leaf 0.2), laserMapping.cpp:1815-1821 and :2160-2166 (lineRes /
// planeRes).  Restated from PCL 1.10 filters/impl/voxel_grid.hpp:
// bbox-relative grid, sort by linear voxel id, centroid of all fields (xyz and intensity) in fp32,
// output in ascending voxel id, "leaf too small" fallback = copy input.
#include "orc_api.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>

namespace orc {

namespace {
struct cloud_point_index_idx {
    unsigned int idx;
    unsigned int cloud_point_index;
    bool operator<(const cloud_point_index_idx& p) const { return idx < p.idx; }
};
}  // namespace

void voxel_grid(const std::vector<P4>& in, float leaf, bool stable, std::vector<P4>& out)
{
    out.clear();
    if (in.empty()) return;  // PCL: empty input -> empty output
    const float inv = 1.0f / leaf;  // inverse_leaf_size_ = Array4f::Ones() / leaf_size_
    float mn[3] = {std::numeric_limits<float>::max(), std::numeric_limits<float>::max(), std::numeric_limits<float>::max()};
    float mx[3] = {-mn[0], -mn[1], -mn[2]};
    for (const P4& p : in) {  // getMinMax3D
        mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
        mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
    }
    const int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1;
    const int64_t dy = (int64_t)((mx[1] - mn[1]) * inv) + 1;
    const int64_t dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
    if (dx * dy * dz > (int64_t)std::numeric_limits<int32_t>::max()) {  // "Leaf size is too small"
        out = in;
        return;
    }
    int min_b[3], max_b[3], div_b[3];
    for (int a = 0; a < 3; ++a) {
        min_b[a] = (int)std::floor(mn[a] * inv);
        max_b[a] = (int)std::floor(mx[a] * inv);
        div_b[a] = max_b[a] - min_b[a] + 1;
    }
    const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};

    std::vector<cloud_point_index_idx> index_vector;
    index_vector.reserve(in.size());
    for (unsigned int it = 0; it < in.size(); ++it) {
        const int ijk0 = (int)(std::floor(in[it].x * inv) - (float)min_b[0]);
        const int ijk1 = (int)(std::floor(in[it].y * inv) - (float)min_b[1]);
        const int ijk2 = (int)(std::floor(in[it].z * inv) - (float)min_b[2]);
        const int idx = ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2];
        index_vector.push_back({(unsigned int)idx, it});
    }
    if (stable)
        std::stable_sort(index_vector.begin(), index_vector.end());
    else
        std::sort(index_vector.begin(), index_vector.end());  // as PCL: order within a voxel unspecified

    size_t index = 0;
    while (index < index_vector.size()) {
        size_t i = index + 1;
        while (i < index_vector.size() && index_vector[i].idx == index_vector[index].idx) ++i;
        // CentroidPoint<PointXYZI>: AccumulatorXYZ (Vector3f sum) + AccumulatorIntensity (float sum)
        float sx = 0.f, sy = 0.f, sz = 0.f, si = 0.f;
        for (size_t li = index; li < i; ++li) {
            const P4& p = in[index_vector[li].cloud_point_index];
            sx += p.x; sy += p.y; sz += p.z; si += p.i;
        }
        const float n = (float)(i - index);
        out.push_back({sx / n, sy / n, sz / n, si / n});
        index = i;
    }
}

}  // namespace orc
