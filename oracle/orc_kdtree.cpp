// ORACLE (test infrastructure) — pcl::KdTreeFLANN<PointXYZI> restated as an exact k-d tree.
//
// PCL/FLANN are un-vendored.  Restated behaviour: setInputCloud copies the
// xyz of every point and builds a single k-d tree (FLANN KDTreeSingleIndexParams(15): leaves of <= 15
// points, data reordered); nearestKSearch is an exact (eps = 0), ascending-sorted k-NN with the fp32
// `L2_Simple` functor: result = 0; for x,y,z: diff = a - b; result += diff * diff.
// Ties: FLANN's order is unspecified; the rule here (and on the GPU) is lowest target index first.
#include "orc_api.h"

#include <algorithm>
#include <cmath>

namespace orc {

void KdTree::build(const std::vector<P4>& pts)
{
    n_ = (int)pts.size();
    nodes_.clear();
    order_.resize(n_);
    xyz_.resize((size_t)n_ * 3);
    if (n_ == 0) return;  // PCL: "Cannot create a KDTree with an empty input cloud!"
    for (int i = 0; i < n_; ++i) order_[i] = i;
    std::vector<float> raw((size_t)n_ * 3);
    for (int i = 0; i < n_; ++i) { raw[3 * i] = pts[i].x; raw[3 * i + 1] = pts[i].y; raw[3 * i + 2] = pts[i].z; }
    xyz_.swap(raw);  // build on the original order, then reorder
    nodes_.reserve((size_t)n_ / 4 + 16);
    build_rec(0, n_);
    std::vector<float> re((size_t)n_ * 3);
    for (int i = 0; i < n_; ++i) for (int a = 0; a < 3; ++a) re[3 * i + a] = xyz_[3 * order_[i] + a];
    xyz_.swap(re);
}

int KdTree::build_rec(int begin, int end)
{
    const int id = (int)nodes_.size();
    nodes_.push_back(Node());
    Node nd;
    nd.left = nd.right = -1;
    nd.begin = begin;
    nd.end = end;
    for (int a = 0; a < 3; ++a) { nd.lo[a] = INFINITY; nd.hi[a] = -INFINITY; }
    for (int i = begin; i < end; ++i)
        for (int a = 0; a < 3; ++a) {
            const float v = xyz_[3 * order_[i] + a];
            nd.lo[a] = std::min(nd.lo[a], v);
            nd.hi[a] = std::max(nd.hi[a], v);
        }
    if (end - begin > 15) {
        int dim = 0;
        float span = nd.hi[0] - nd.lo[0];
        for (int a = 1; a < 3; ++a)
            if (nd.hi[a] - nd.lo[a] > span) { span = nd.hi[a] - nd.lo[a]; dim = a; }
        if (span > 0.f) {
            const float cut = 0.5f * (nd.lo[dim] + nd.hi[dim]);
            int mid = (int)(std::partition(order_.begin() + begin, order_.begin() + end,
                                           [&](int p) { return xyz_[3 * p + dim] < cut; }) - order_.begin());
            if (mid == begin || mid == end) {  // degenerate split: fall back to the median
                mid = (begin + end) / 2;
                std::nth_element(order_.begin() + begin, order_.begin() + mid, order_.begin() + end,
                                 [&](int p, int q) { return xyz_[3 * p + dim] < xyz_[3 * q + dim]; });
            }
            nd.left = build_rec(begin, mid);
            nd.right = build_rec(mid, end);
        }
    }
    nodes_[id] = nd;
    return id;
}

static inline double box_lb(const float lo[3], const float hi[3], const float q[3])
{
    double s = 0.0;
    for (int a = 0; a < 3; ++a) {
        double d = 0.0;
        if (q[a] < lo[a]) d = (double)lo[a] - (double)q[a];
        else if (q[a] > hi[a]) d = (double)q[a] - (double)hi[a];
        s += d * d;
    }
    return s;
}

void KdTree::search(int node, const float q[3], int k, int* idx, float* d2, int& found) const
{
    const Node& nd = nodes_[node];
    if (nd.left < 0) {
        for (int i = nd.begin; i < nd.end; ++i) {
            float r = 0.f;  // L2_Simple, fp32, x then y then z
            for (int a = 0; a < 3; ++a) { const float diff = q[a] - xyz_[3 * i + a]; r += diff * diff; }
            const int oi = order_[i];
            if (found == k && !(r < d2[k - 1] || (r == d2[k - 1] && oi < idx[k - 1]))) continue;
            int pos = found < k ? found : k - 1;
            while (pos > 0 && (r < d2[pos - 1] || (r == d2[pos - 1] && oi < idx[pos - 1]))) {
                d2[pos] = d2[pos - 1];
                idx[pos] = idx[pos - 1];
                --pos;
            }
            d2[pos] = r;
            idx[pos] = oi;
            if (found < k) ++found;
        }
        return;
    }
    const Node& L = nodes_[nd.left];
    const Node& R = nodes_[nd.right];
    const double lbL = box_lb(L.lo, L.hi, q), lbR = box_lb(R.lo, R.hi, q);
    const int first = lbL <= lbR ? nd.left : nd.right, second = lbL <= lbR ? nd.right : nd.left;
    const double lbF = std::min(lbL, lbR), lbS = std::max(lbL, lbR);
    // conservative pruning: fp32 distances are within a few ulp of the true ones
    if (found < k || lbF * (1.0 - 1e-6) <= (double)d2[k - 1]) search(first, q, k, idx, d2, found);
    if (found < k || lbS * (1.0 - 1e-6) <= (double)d2[k - 1]) search(second, q, k, idx, d2, found);
}

int KdTree::knn(const float q[3], int k, int* idx, float* d2) const
{
    if (n_ == 0) return 0;
    if (k > n_) k = n_;
    int found = 0;
    search(0, q, k, idx, d2, found);
    return found;
}

}  // namespace orc
