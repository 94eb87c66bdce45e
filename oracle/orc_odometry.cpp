// ORACLE (test infrastructure) — laserOdometry.cpp restated: TransformToStart (LO:77-95), Distance
// (LO:153-162), graph_based_correspondence_vote_simple (LO:165-342) and one pass of the main loop
// body (LO:425-896) for one synchronized set of feature clouds.
//
// fp conventions: association distances are fp32 expressions promoted to double on assignment
// (LO:514-519 etc.); TransformToStart computes in fp64 and stores fp32; the vote is fp32 sqrt/exp.
#include "orc_api.h"
#include "orc_jet.h"

#include <algorithm>
#include <cmath>

namespace orc {

namespace {

const double DISTANCE_SQ_THRESHOLD = 25;  // LO:29
const double NEARBY_SCAN = 2.5;           // LO:30

const double SCAN_PERIOD = 0.1;           // LO:28

// interpolation ratio of a point (LO:81-84, 569-573, 739-743): DISTORTION 0 (the reference build) -> 1.0;
// DISTORTION 1 -> fraction of the intensity (= 0.1 * relTime, SR:208) / SCAN_PERIOD
inline double point_s(const P4& p, int distortion)
{
    if (distortion) return (p.i - int(p.i)) / SCAN_PERIOD;   // float - int -> float, then / double
    return 1.0;
}

// LO:77-95
inline P4 transform_to_start(const P4& pi, const double para_q[4], const double para_t[3], double s)
{
    const Quat<double> q_last_curr{para_q[0], para_q[1], para_q[2], para_q[3]};
    const Quat<double> q_point_last = identity_slerp(s, q_last_curr);
    const V3<double> t_point_last{s * para_t[0], s * para_t[1], s * para_t[2]};
    const V3<double> point{pi.x, pi.y, pi.z};
    const V3<double> un_point = rotate(q_point_last, point) + t_point_last;
    return {(float)un_point.x, (float)un_point.y, (float)un_point.z, pi.i};
}

inline float Distance(const P4& a, const P4& b)  // LO:153-162
{
    float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    return std::sqrt(dx * dx + dy * dy + dz * dz);
}

inline double sqdist(const P4& p, const P4& q)  // LO:514-519: fp32 expression, widened on assignment
{
    return (p.x - q.x) * (p.x - q.x) + (p.y - q.y) * (p.y - q.y) + (p.z - q.z) * (p.z - q.z);
}

struct compare_score {  // common.h:50-52
    bool operator()(VertexVote const& ob1, VertexVote const& ob2) { return ob1.score > ob2.score; }
};

}  // namespace

void graph_vote_simple(const std::vector<CorreMatch>& correspondences, bool corner_case, std::vector<VertexVote>& selected_idx,
                       std::vector<float>* votes_out)
{
    const int cor_size_all = (int)correspondences.size();
    const int number_of_region = corner_case ? 5 : 10;  // LO:179-188
    const float score_threshold = 0.96;
    if (votes_out) votes_out->assign(cor_size_all, 0.f);
    for (int num_region = 0; num_region < number_of_region; num_region++) {  // LO:193
        const int initial_pos = cor_size_all / number_of_region * (num_region);
        const int end_pos = (num_region == number_of_region - 1) ? cor_size_all : cor_size_all / number_of_region * (num_region + 1);
        const int cor_size = end_pos - initial_pos;
        const CorreMatch* sel = correspondences.data() + initial_pos;
        const float resolution = 1;
        std::vector<VertexVote> vote_record(cor_size, VertexVote{0, 0.0f});
        for (int i = 0; i < cor_size; i++) {  // LO:228-252 (the compatibility matrix is a dead store, LO:240-241)
            vote_record[i].index = i;
            for (int j = i + 1; j < cor_size; j++) {
                const float s1 = Distance(sel[i].src, sel[j].src);
                const float s2 = Distance(sel[i].tgt, sel[j].tgt);
                const float dis_gap = std::abs(s1 - s2);
                const float score = std::exp(-(dis_gap * dis_gap) / (resolution * resolution));
                if (score < score_threshold) {
                    vote_record[j].score += 1;
                    vote_record[i].score += 1;
                }
            }
        }
        if (votes_out) for (int i = 0; i < cor_size; ++i) (*votes_out)[initial_pos + i] = vote_record[i].score;
        std::sort(vote_record.begin(), vote_record.end(), compare_score());  // LO:255
        // LO:257-330: the corner and plane branches are textually identical
        const float selected_ratio = 0.90;
        const float num_selected = selected_ratio * cor_size;
        const float selected_count_ratio = 1;
        const int donot_num_selected = (1 - selected_count_ratio) * cor_size;
        for (int i = cor_size - 1; i >= 0; i--) {
            if (i >= donot_num_selected) {
                VertexVote obj;
                if (sel[vote_record[i].index].index > (int)correspondences.size()) continue;
                obj.index = sel[vote_record[i].index].index;
                if (vote_record[i].score > num_selected) {
                    obj.score = 0;
                    break;
                } else if (vote_record[i].score <= 50) {
                    obj.score = 5.0;
                } else {
                    obj.score = 1;
                }
                selected_idx.push_back(obj);
            }
        }
    }
}

// graph_based_correspondence_vote_partial (laserMapping.cpp:321-834; graph_construction_partial LM:261-318) - the
// paper-style scoring.  DEAD in the reference: the only call is commented out (LO:622), and the function lives in the
// mapping node.  Restated for the optional mode cfg.vote_mode = 1, where it scores the odometry's plane correspondences
// in place of vote_simple (beyond-reference behaviour).  Per region of m correspondences:
//   G(i,j) = expf(-gap^2); neighbours N(i) = {j : G(i,j) > 0.95}; first pass: s_i = mean over neighbour pairs (a,b) with
//   G(a,b) != 0 of cbrt(G(i,a) G(i,b) G(a,b)) (pairs with G(a,b) = 0 count as 0); threshold = min(sum_i num_i / sum_i den_i,
//   mean_i s_i); neighbours with s < threshold are pruned; final score = 0.1 * mean G(a,i) + 0.9 * tight, where - as
//   written, std::pow(x, 1/3) with the INTEGER quotient 1/3 = 0 - every surviving neighbour pair with G(a,b) != 0
//   contributes 1, divided by the integer d (d - 2) / 2; both parts are 0 unless d > 2.  Selected = score != 0, in
//   descending score order, weight = score.
void graph_vote_partial(const std::vector<CorreMatch>& correspondences, bool corner_case, std::vector<VertexVote>& selected_idx)
{
    (void)corner_case;   // LM:716-787: the two branches are textually identical (selected_ratio = 1)
    const int cor_size_all = (int)correspondences.size();
    const int number_of_region = 10;   // LM:327
    for (int num_region = 0; num_region < number_of_region; num_region++) {
        const int initial_pos = cor_size_all / number_of_region * (num_region);
        const int end_pos = (num_region == number_of_region - 1) ? cor_size_all : cor_size_all / number_of_region * (num_region + 1);
        const int cor_size = end_pos - initial_pos;
        const CorreMatch* sel = correspondences.data() + initial_pos;
        // graph_construction_partial, LM:261-318
        std::vector<float> G((size_t)cor_size * cor_size, 0.f);
        double gnorm = 0.0;
        for (int i = 0; i < cor_size; i++)
            for (int j = i + 1; j < cor_size; j++) {
                const float s1 = Distance(sel[i].src, sel[j].src);
                const float s2 = Distance(sel[i].tgt, sel[j].tgt);
                const float dis_gap = std::abs(s1 - s2);
                const float resolution = 1;
                const float score = std::exp(-(dis_gap * dis_gap) / (resolution * resolution));
                G[(size_t)i * cor_size + j] = score;
                G[(size_t)j * cor_size + i] = score;
                gnorm += (double)score * score;
            }
        if (gnorm == 0) continue;   // LM:399-403 "Graph is not connected!"
        auto g = [&](int i, int j) { return G[(size_t)i * cor_size + j]; };
        std::vector<std::vector<int>> conn(cor_size);
        std::vector<int> degree(cor_size, 0);
        for (int i = 0; i < cor_size; i++)   // LM:407-430
            for (int j = 0; j < cor_size; j++)
                if (i != j && g(i, j) > 0.95) { degree[i]++; conn[i].push_back(j); }
        std::vector<VertexVote> neighbor_voter;
        float filter_param_numerator_a = 0, filter_param_denominator_a = 0, filter_param_b = 0;
        for (int i = 0; i < cor_size; i++) {   // LM:445-500
            VertexVote ob;
            std::vector<float> weight_s((size_t)(degree[i] * (degree[i] - 1) * 0.5), 0);
            for (int j = 0; j < degree[i]; j++) {
                const int a = conn[i][j];
                int count_b = 0;
                for (int k = j + 1; k < degree[i]; k++) {
                    const int b = conn[i][k];
                    if (g(a, b)) weight_s[int(j * (2 * degree[i] - 1 - j) / 2 + count_b)] = std::pow(g(i, a) * g(i, b) * g(a, b), 1.0 / 3);
                    count_b++;
                }
            }
            if (degree[i] > 1) {
                double acc = 0.0;
                for (float w : weight_s) acc += w;   // std::accumulate(..., 0.0)
                const float numerator = acc;
                const float denominator = degree[i] * (degree[i] - 1) * 0.5;
                filter_param_numerator_a += numerator;
                filter_param_denominator_a += denominator;
                ob.index = i;
                ob.score = numerator / denominator;
            } else {
                ob.index = i;
                ob.score = 0;
            }
            neighbor_voter.push_back(ob);
            filter_param_b += ob.score;
        }
        const float filter_param_a = filter_param_numerator_a / filter_param_denominator_a;
        filter_param_b = filter_param_b / neighbor_voter.size();
        const float threshold_param = std::min(filter_param_a, filter_param_b);
        for (int i = 0; i < cor_size; i++) {   // LM:560-580 prune
            std::vector<int> index_pruned;
            for (int idx : conn[i])
                if (neighbor_voter[idx].score >= threshold_param) index_pruned.push_back(idx);
            conn[i] = index_pruned;
            degree[i] = (int)index_pruned.size();
        }
        const float weight_balance = 0.9;
        std::vector<VertexVote> voter;
        for (int i = 0; i < cor_size; i++) {   // LM:600-690
            float score_all = 0;
            const size_t d = conn[i].size();
            std::vector<float> looser_score(d, 0);
            std::vector<float> tight_score((size_t)(d * (d - 1) * 0.5), 0);
            float tight_score_sum = 0, looser_score_sum = 0;
            if (d > 2) {
                for (size_t j = 0; j < d; j++) {
                    const int idx_a = conn[i][j];
                    looser_score[j] = g(idx_a, i);
                    int count_b = 0;
                    for (size_t k = j + 1; k < d; k++) {
                        const int idx_b = conn[i][k];
                        if (g(idx_a, idx_b)) tight_score[int(j * (2 * d - 1 - j) / 2 + count_b)] = std::pow(g(idx_a, idx_b) * g(idx_a, i) * g(idx_b, i), 1 / 3);
                        count_b++;
                    }
                }
                double acc = 0.0;
                for (float w : tight_score) acc += w;
                tight_score_sum = acc;
                tight_score_sum /= (degree[i] * (degree[i] - 2) / 2);
            }
            if (degree[i] != 0) {
                double acc = 0.0;
                for (float w : looser_score) acc += w;
                looser_score_sum = acc;
                looser_score_sum = looser_score_sum / degree[i];
            }
            score_all = (1 - weight_balance) * looser_score_sum + weight_balance * tight_score_sum;
            VertexVote obj;
            obj.index = i;
            obj.score = score_all;
            voter.push_back(obj);
        }
        std::vector<VertexVote> voter_ordered(voter.begin(), voter.end());
        std::sort(voter_ordered.begin(), voter_ordered.end(), compare_score());   // LM:700
        const float selected_ratio = 1;
        const int num_selected = selected_ratio * cor_size;
        for (int i = 0; i < cor_size; i++) {   // LM:716-787
            if (i < num_selected) {
                VertexVote obj;
                if (sel[voter_ordered[i].index].index > (int)correspondences.size()) continue;
                obj.index = sel[voter_ordered[i].index].index;
                obj.score = voter_ordered[i].score;
                if (obj.score != 0) selected_idx.push_back(obj);
            } else {
                break;
            }
        }
    }
}

void Odometry::step(const std::vector<P4>& cornerPointsSharp, const std::vector<P4>& cornerPointsLessSharp,
                    const std::vector<P4>& surfPointsFlat, const std::vector<P4>& surfPointsLessFlat)
{
    last_stats.clear();
    const bool was_inited = systemInited;
    if (!systemInited) {  // LO:427-431
        systemInited = true;
    } else {
        const int cornerPointsSharpNum = (int)cornerPointsSharp.size();
        const int surfPointsFlatNum = (int)surfPointsFlat.size();
        const std::vector<P4>& laserCloudCornerLast = cornerLast;
        const std::vector<P4>& laserCloudSurfLast = surfLast;
        for (size_t opti_counter = 0; opti_counter < 3; ++opti_counter) {  // LO:439
            OdomIterStats st{};
            std::vector<ResidualBlock> problem;
            last_corner_assoc.clear();
            last_plane_assoc.clear();
            int idx1[1];
            float d1[1];

            // LO:491-620 corner correspondences
            for (int i = 0; i < cornerPointsSharpNum; ++i) {
                const P4 pointSel = transform_to_start(cornerPointsSharp[i], para_q, para_t, point_s(cornerPointsSharp[i], cfg.distortion));
                const float qf[3] = {pointSel.x, pointSel.y, pointSel.z};
                if (kdCorner.knn(qf, 1, idx1, d1) < 1) continue;  // empty tree: PCL returns 0 neighbours
                int closestPointInd = -1, minPointInd2 = -1;
                if (d1[0] < DISTANCE_SQ_THRESHOLD) {
                    closestPointInd = idx1[0];
                    const int closestPointScanID = int(laserCloudCornerLast[closestPointInd].i);
                    double minPointSqDis2 = DISTANCE_SQ_THRESHOLD;
                    for (int j = closestPointInd + 1; j < (int)laserCloudCornerLast.size(); ++j) {  // increasing scan line
                        if (int(laserCloudCornerLast[j].i) <= closestPointScanID) continue;
                        if (int(laserCloudCornerLast[j].i) > (closestPointScanID + NEARBY_SCAN)) break;
                        const double pointSqDis = sqdist(laserCloudCornerLast[j], pointSel);
                        if (pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
                    }
                    for (int j = closestPointInd - 1; j >= 0; --j) {  // decreasing scan line
                        if (int(laserCloudCornerLast[j].i) >= closestPointScanID) continue;
                        if (int(laserCloudCornerLast[j].i) < (closestPointScanID - NEARBY_SCAN)) break;
                        const double pointSqDis = sqdist(laserCloudCornerLast[j], pointSel);
                        if (pointSqDis < minPointSqDis2) { minPointSqDis2 = pointSqDis; minPointInd2 = j; }
                    }
                }
                if (minPointInd2 >= 0) {  // LO:556-618
                    const double cp[3] = {cornerPointsSharp[i].x, cornerPointsSharp[i].y, cornerPointsSharp[i].z};
                    const double a[3] = {laserCloudCornerLast[closestPointInd].x, laserCloudCornerLast[closestPointInd].y, laserCloudCornerLast[closestPointInd].z};
                    const double b[3] = {laserCloudCornerLast[minPointInd2].x, laserCloudCornerLast[minPointInd2].y, laserCloudCornerLast[minPointInd2].z};
                    problem.push_back(make_edge(cp, a, b, point_s(cornerPointsSharp[i], cfg.distortion)));   // LO:569-573
                    last_corner_assoc.push_back({i, closestPointInd, minPointInd2});
                    st.corner_corr++;
                }
            }

            // LO:653-793 plane correspondences
            std::vector<CorreMatch> correspondences;
            std::vector<ResidualBlock> plane_blocks;  // weight 1; re-weighted after the vote
            std::vector<std::array<double, 12>> plane_pts;
            std::vector<double> plane_s;   // LO:739-743
            int index = 0;
            for (int i = 0; i < surfPointsFlatNum; ++i) {
                const P4 pointSel = transform_to_start(surfPointsFlat[i], para_q, para_t, point_s(surfPointsFlat[i], cfg.distortion));
                const float qf[3] = {pointSel.x, pointSel.y, pointSel.z};
                if (kdSurf.knn(qf, 1, idx1, d1) < 1) continue;
                int closestPointInd = -1, minPointInd2 = -1, minPointInd3 = -1;
                if (d1[0] < DISTANCE_SQ_THRESHOLD) {
                    closestPointInd = idx1[0];
                    const int closestPointScanID = int(laserCloudSurfLast[closestPointInd].i);
                    double minPointSqDis2 = DISTANCE_SQ_THRESHOLD, minPointSqDis3 = DISTANCE_SQ_THRESHOLD;
                    for (int j = closestPointInd + 1; j < (int)laserCloudSurfLast.size(); ++j) {
                        if (int(laserCloudSurfLast[j].i) > (closestPointScanID + NEARBY_SCAN)) break;
                        const double pointSqDis = sqdist(laserCloudSurfLast[j], pointSel);
                        if (int(laserCloudSurfLast[j].i) <= closestPointScanID && pointSqDis < minPointSqDis2) {
                            minPointSqDis2 = pointSqDis; minPointInd2 = j;
                        } else if (int(laserCloudSurfLast[j].i) > closestPointScanID && pointSqDis < minPointSqDis3) {
                            minPointSqDis3 = pointSqDis; minPointInd3 = j;
                        }
                    }
                    for (int j = closestPointInd - 1; j >= 0; --j) {
                        if (int(laserCloudSurfLast[j].i) < (closestPointScanID - NEARBY_SCAN)) break;
                        const double pointSqDis = sqdist(laserCloudSurfLast[j], pointSel);
                        if (int(laserCloudSurfLast[j].i) >= closestPointScanID && pointSqDis < minPointSqDis2) {
                            minPointSqDis2 = pointSqDis; minPointInd2 = j;
                        } else if (int(laserCloudSurfLast[j].i) < closestPointScanID && pointSqDis < minPointSqDis3) {
                            minPointSqDis3 = pointSqDis; minPointInd3 = j;
                        }
                    }
                    if (minPointInd2 >= 0 && minPointInd3 >= 0) {  // LO:723-791
                        const P4& pa = laserCloudSurfLast[closestPointInd];
                        const P4& pb = laserCloudSurfLast[minPointInd2];
                        const P4& pc = laserCloudSurfLast[minPointInd3];
                        plane_pts.push_back({(double)surfPointsFlat[i].x, (double)surfPointsFlat[i].y, (double)surfPointsFlat[i].z,
                                             (double)pa.x, (double)pa.y, (double)pa.z, (double)pb.x, (double)pb.y, (double)pb.z,
                                             (double)pc.x, (double)pc.y, (double)pc.z});
                        plane_s.push_back(point_s(surfPointsFlat[i], cfg.distortion));
                        CorreMatch cor;
                        cor.index = index;
                        cor.src = surfPointsFlat[i];
                        cor.tgt = pa;
                        cor.score = 0;
                        cor.s = 1 / 1.0;
                        correspondences.push_back(cor);
                        index++;
                        last_plane_assoc.push_back({i, closestPointInd, minPointInd2, minPointInd3});
                        if (now_frame <= cfg.graph_from_frame) {  // LO:781-787
                            const auto& pp = plane_pts.back();
                            problem.push_back(make_plane_modify(&pp[0], &pp[3], &pp[6], &pp[9], plane_s.back(), 1));
                            st.plane_selected++;
                        }
                        st.plane_corr++;
                    }
                }
            }
            if (now_frame > cfg.graph_from_frame) {  // LO:794-810
                std::vector<VertexVote> selected_idx;
                if (cfg.vote_mode == 1) graph_vote_partial(correspondences, false, selected_idx);   // optional, beyond-reference (dead code LM:321-834)
                else graph_vote_simple(correspondences, false, selected_idx, nullptr);
                for (size_t i = 0; i < selected_idx.size(); i++) {
                    const auto& pp = plane_pts[selected_idx[i].index];
                    problem.push_back(make_plane_modify(&pp[0], &pp[3], &pp[6], &pp[9], plane_s[selected_idx[i].index], selected_idx[i].score));
                }
                st.plane_selected = (int)selected_idx.size();
            }
            solve(problem, para_q, para_t, &st.solve, 4, true);  // LO:819-825
            last_stats.push_back(st);
        }
        // LO:830-831
        const Quat<double> qw{q_w_curr[0], q_w_curr[1], q_w_curr[2], q_w_curr[3]};
        const Quat<double> ql{para_q[0], para_q[1], para_q[2], para_q[3]};
        const V3<double> r = rotate(qw, V3<double>{para_t[0], para_t[1], para_t[2]});
        t_w_curr[0] = t_w_curr[0] + r.x;
        t_w_curr[1] = t_w_curr[1] + r.y;
        t_w_curr[2] = t_w_curr[2] + r.z;
        const Quat<double> qn = qmul(qw, ql);
        q_w_curr[0] = qn.x; q_w_curr[1] = qn.y; q_w_curr[2] = qn.z; q_w_curr[3] = qn.w;
    }
    // LO:882-896: swap in the less-sharp / less-flat clouds and rebuild both kd-trees
    cornerLast = cornerPointsLessSharp;
    surfLast = surfPointsLessFlat;
    // LO:861-880 TransformToEnd of the clouds about to become *Last.  The reference guards the block with a literal
    // `if (0)` even when DISTORTION is 1; cfg.distortion == 2 is DISTORTION 1 with that block enabled (the LOAM-family
    // de-skew in full).  Runs only for frames that went through the solve (the block sits inside the else branch).
    if (cfg.distortion == 2 && was_inited) {
        const Quat<double> ql{para_q[0], para_q[1], para_q[2], para_q[3]};
        const Quat<double> qi = qinverse(ql);
        auto to_end = [&](P4& p) {  // LO:98-114
            const P4 un = transform_to_start(p, para_q, para_t, point_s(p, 1));   // stored as fp32 (pcl::PointXYZI un_point_tmp)
            const V3<double> d{(double)un.x - para_t[0], (double)un.y - para_t[1], (double)un.z - para_t[2]};
            const V3<double> e = rotate(qi, d);
            p = P4{(float)e.x, (float)e.y, (float)e.z, (float)int(p.i)};
        };
        for (P4& p : cornerLast) to_end(p);
        for (P4& p : surfLast) to_end(p);
    }
    kdCorner.build(cornerLast);
    kdSurf.build(surfLast);
    now_frame++;  // LO:926
}

}  // namespace orc
