// ORACLE (test infrastructure) — laserMapping.cpp `process()` restated for one frame: LM:1581 (initial
// guess), LM:1584-1811 (cube shift + 5x5x3 local map), LM:1814-1822 (voxel DS of the incoming clouds),
// LM:1826-2100 (kd-tree build, 2 x { 5-NN + line / plane fit + ceres::Solve }), LM:2101 (transformUpdate),
// LM:2104-2168 (map insert + per-cube voxel DS).  Publishing and the trajectory file are out of scope.
//
// Eigen restatements: SelfAdjointEigenSolver<Matrix3d> -> cyclic Jacobi (eigenvalues
// ascending); colPivHouseholderQr().solve on the 5x3 system -> column-pivoted Householder QR with
// Eigen's default rank threshold.
#include "orc_api.h"
#include "orc_jet.h"

#include <algorithm>
#include <cmath>

namespace orc {

void sym_eig3(const double Ain[9], double evals[3], double evecs[9])
{
    double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i][j] = Ain[i * 3 + j];
    for (int sweep = 0; sweep < 60; ++sweep) {
        const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
        const double diag = A[0][0] * A[0][0] + A[1][1] * A[1][1] + A[2][2] * A[2][2];
        if (off <= 1e-32 * diag || off == 0.0) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {  // A <- A J
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {  // A <- J^T A
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    int ord[3] = {0, 1, 2};
    std::sort(ord, ord + 3, [&](int a, int b) { return A[a][a] < A[b][b]; });
    for (int c = 0; c < 3; ++c) {
        evals[c] = A[ord[c]][ord[c]];
        for (int r = 0; r < 3; ++r) evecs[r * 3 + c] = V[r][ord[c]];
    }
}

bool plane_fit5(const double pts[15], double nrm[3])
{
    // min || A n - b ||, A = pts (5x3), b = -1; column-pivoted Householder QR
    double A[5][3], b[5];
    int perm[3] = {0, 1, 2};
    for (int i = 0; i < 5; ++i) { for (int j = 0; j < 3; ++j) A[i][j] = pts[i * 3 + j]; b[i] = -1.0; }
    double colnorm[3];
    for (int j = 0; j < 3; ++j) { colnorm[j] = 0; for (int i = 0; i < 5; ++i) colnorm[j] += A[i][j] * A[i][j]; }
    double maxpivot = 0.0;
    double diag[3] = {0, 0, 0};
    int rank = 3;
    for (int k = 0; k < 3; ++k) {
        int piv = k;
        double best = -1.0;
        for (int j = k; j < 3; ++j) {
            double s = 0;
            for (int i = k; i < 5; ++i) s += A[i][j] * A[i][j];
            if (s > best) { best = s; piv = j; }
        }
        if (piv != k) {
            for (int i = 0; i < 5; ++i) std::swap(A[i][k], A[i][piv]);
            std::swap(perm[k], perm[piv]);
        }
        const double nrmk = std::sqrt(best);
        if (nrmk == 0.0) { rank = k; break; }
        const double alpha = A[k][k] > 0 ? -nrmk : nrmk;
        const double v0 = A[k][k] - alpha;
        double vtv = v0 * v0;
        for (int i = k + 1; i < 5; ++i) vtv += A[i][k] * A[i][k];
        if (vtv > 0.0) {
            for (int j = k + 1; j < 3; ++j) {
                double s = v0 * A[k][j];
                for (int i = k + 1; i < 5; ++i) s += A[i][k] * A[i][j];
                s = 2.0 * s / vtv;
                A[k][j] -= s * v0;
                for (int i = k + 1; i < 5; ++i) A[i][j] -= s * A[i][k];
            }
            double s = v0 * b[k];
            for (int i = k + 1; i < 5; ++i) s += A[i][k] * b[i];
            s = 2.0 * s / vtv;
            b[k] -= s * v0;
            for (int i = k + 1; i < 5; ++i) b[i] -= s * A[i][k];
        }
        A[k][k] = alpha;
        diag[k] = std::fabs(alpha);
        if (diag[k] > maxpivot) maxpivot = diag[k];
    }
    // Eigen: rank = #pivots with |pivot| > epsilon * diagonalSize * maxpivot; the rest of the solution is zero
    const double thr = 2.220446049250313e-16 * 3.0 * maxpivot;
    int r = 0;
    for (int k = 0; k < rank; ++k) if (diag[k] > thr) ++r;
    double y[3] = {0, 0, 0};
    for (int k = r - 1; k >= 0; --k) {
        double s = b[k];
        for (int j = k + 1; j < r; ++j) s -= A[k][j] * y[j];
        y[k] = s / A[k][k];
    }
    for (int k = 0; k < 3; ++k) nrm[perm[k]] = y[k];
    (void)colnorm;
    return r == 3;
}

namespace {

inline P4 associate_to_map(const P4& pi, const double* par)  // LM:125-134
{
    const Quat<double> q{par[0], par[1], par[2], par[3]};
    const V3<double> pw = rotate(q, V3<double>{pi.x, pi.y, pi.z}) + V3<double>{par[4], par[5], par[6]};
    return {(float)pw.x, (float)pw.y, (float)pw.z, pi.i};
}

inline void cube_of(const P4& p, int cenW, int cenH, int cenD, int& I, int& J, int& K)  // LM:2109-2118
{
    I = int((p.x + 25.0) / 50.0) + cenW;
    J = int((p.y + 25.0) / 50.0) + cenH;
    K = int((p.z + 25.0) / 50.0) + cenD;
    if (p.x + 25.0 < 0) I--;
    if (p.y + 25.0 < 0) J--;
    if (p.z + 25.0 < 0) K--;
}

}  // namespace

void Mapping::insert_map_points(const std::vector<P4>& corner, const std::vector<P4>& surf)
{
    for (int pass = 0; pass < 2; ++pass) {
        const std::vector<P4>& src = pass == 0 ? corner : surf;
        std::vector<std::vector<P4>>& dst = pass == 0 ? cornerArray : surfArray;
        for (const P4& p : src) {
            int I, J, K;
            cube_of(p, cenW, cenH, cenD, I, J, K);
            if (I >= 0 && I < W && J >= 0 && J < H && K >= 0 && K < D) dst[I + W * J + W * H * K].push_back(p);
        }
    }
}

// graph_based_correspondence_vote_simple as laserMapping.cpp carries it (LM:836-1027; Distance LM:250-259): 20 contiguous
// regions, a vote when  std::exp(-(gap * gap) / (1 * 1)) < 0.95  (float score against a double literal), and in the
// corner_case branch - the only one this copy has, and the one the commented call passes - every correspondence with
// fewer votes than 0.75 * region size is selected with score 1.0, walking the descending sort from its end.
void graph_vote_simple_mapping(const std::vector<CorreMatch>& correspondences, bool corner_case, std::vector<VertexVote>& selected_idx)
{
    struct by_score { bool operator()(VertexVote const& a, VertexVote const& b) { return a.score > b.score; } };  // common.h:50-52
    auto Distance = [](const P4& a, const P4& b) { float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z; return std::sqrt(dx * dx + dy * dy + dz * dz); };
    const int cor_size_all = (int)correspondences.size();
    const int number_of_region = 20;
    for (int num_region = 0; num_region < number_of_region; num_region++) {
        const int initial_pos = cor_size_all / number_of_region * (num_region);
        const int end_pos = (num_region == number_of_region - 1) ? cor_size_all : cor_size_all / number_of_region * (num_region + 1);
        const int cor_size = end_pos - initial_pos;
        const CorreMatch* sel = correspondences.data() + initial_pos;
        const float resolution = 1;
        std::vector<VertexVote> vote_record(cor_size, VertexVote{0, 0.0f});
        for (int i = 0; i < cor_size; i++) {
            vote_record[i].index = i;
            for (int j = i + 1; j < cor_size; j++) {
                const float s1 = Distance(sel[i].src, sel[j].src);
                const float s2 = Distance(sel[i].tgt, sel[j].tgt);
                const float dis_gap = std::abs(s1 - s2);
                const float score = std::exp(-(dis_gap * dis_gap) / (resolution * resolution));
                if (score < 0.95) {
                    vote_record[j].score += 1;
                    vote_record[i].score += 1;
                }
            }
        }
        std::sort(vote_record.begin(), vote_record.end(), by_score());
        if (corner_case) {
            const float selected_ratio = 0.75;
            const float num_selected = selected_ratio * cor_size;
            for (int i = cor_size - 1; i >= 0; i--) {
                if (vote_record[i].score < num_selected) {
                    VertexVote obj;
                    if (sel[vote_record[i].index].index > (int)correspondences.size()) continue;
                    obj.index = sel[vote_record[i].index].index;
                    obj.score = 1.0;
                    selected_idx.push_back(obj);
                } else {
                    break;
                }
            }
        }
    }
}

int Mapping::step(const std::vector<P4>& laserCloudCornerLast, const std::vector<P4>& laserCloudSurfLast,
                  const double q_wodom_curr[4], const double t_wodom_curr[3])
{
    last_solves.clear();
    double* q_w_curr = parameters;
    double* t_w_curr = parameters + 4;
    const Quat<double> qwo{q_wodom_curr[0], q_wodom_curr[1], q_wodom_curr[2], q_wodom_curr[3]};
    const V3<double> two{t_wodom_curr[0], t_wodom_curr[1], t_wodom_curr[2]};
    {  // transformAssociateToMap, LM:113-117
        const Quat<double> qmw{q_wmap_wodom[0], q_wmap_wodom[1], q_wmap_wodom[2], q_wmap_wodom[3]};
        const Quat<double> q = qmul(qmw, qwo);
        const V3<double> t = rotate(qmw, two) + V3<double>{t_wmap_wodom[0], t_wmap_wodom[1], t_wmap_wodom[2]};
        q_w_curr[0] = q.x; q_w_curr[1] = q.y; q_w_curr[2] = q.z; q_w_curr[3] = q.w;
        t_w_curr[0] = t.x; t_w_curr[1] = t.y; t_w_curr[2] = t.z;
    }

    // LM:1584-1594
    int centerCubeI = int((t_w_curr[0] + 25.0) / 50.0) + cenW;
    int centerCubeJ = int((t_w_curr[1] + 25.0) / 50.0) + cenH;
    int centerCubeK = int((t_w_curr[2] + 25.0) / 50.0) + cenD;
    if (t_w_curr[0] + 25.0 < 0) centerCubeI--;
    if (t_w_curr[1] + 25.0 < 0) centerCubeJ--;
    if (t_w_curr[2] + 25.0 < 0) centerCubeK--;

    // LM:1596-1779: slide the 21x21x11 cube arrays so the centre cube stays >= 3 cubes from every border.
    // One shift along an axis moves every cube one slot, recycles the slot that falls off the far end
    // (cleared) at the near end.  dir = +1: contents move towards higher index (centre was too low).
    auto shift = [&](int axis, int dir) {
        const int n[3] = {W, H, D};
        const int stride[3] = {1, W, W * H};
        const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
        for (int u = 0; u < n[a1]; ++u)
            for (int v = 0; v < n[a2]; ++v) {
                const int base = u * stride[a1] + v * stride[a2];
                for (std::vector<std::vector<P4>>* arr : {&cornerArray, &surfArray}) {
                    if (dir > 0) {
                        std::vector<P4> recycled;
                        recycled.swap((*arr)[base + (n[axis] - 1) * stride[axis]]);
                        for (int i = n[axis] - 1; i >= 1; --i) (*arr)[base + i * stride[axis]].swap((*arr)[base + (i - 1) * stride[axis]]);
                        recycled.clear();
                        (*arr)[base].swap(recycled);
                    } else {
                        std::vector<P4> recycled;
                        recycled.swap((*arr)[base]);
                        for (int i = 0; i < n[axis] - 1; ++i) (*arr)[base + i * stride[axis]].swap((*arr)[base + (i + 1) * stride[axis]]);
                        recycled.clear();
                        (*arr)[base + (n[axis] - 1) * stride[axis]].swap(recycled);
                    }
                }
            }
    };
    while (centerCubeI < 3) { shift(0, +1); centerCubeI++; cenW++; }
    while (centerCubeI >= W - 3) { shift(0, -1); centerCubeI--; cenW--; }
    while (centerCubeJ < 3) { shift(1, +1); centerCubeJ++; cenH++; }
    while (centerCubeJ >= H - 3) { shift(1, -1); centerCubeJ--; cenH--; }
    while (centerCubeK < 3) { shift(2, +1); centerCubeK++; cenD++; }
    while (centerCubeK >= D - 3) { shift(2, -1); centerCubeK--; cenD--; }

    // LM:1781-1811 local map = 5 x 5 x 3 cubes around the centre
    std::vector<int> validInd;
    for (int i = centerCubeI - 2; i <= centerCubeI + 2; i++)
        for (int j = centerCubeJ - 2; j <= centerCubeJ + 2; j++)
            for (int k = centerCubeK - 1; k <= centerCubeK + 1; k++)
                if (i >= 0 && i < W && j >= 0 && j < H && k >= 0 && k < D) validInd.push_back(i + W * j + W * H * k);
    std::vector<P4> cornerFromMap, surfFromMap;
    for (int ind : validInd) {
        cornerFromMap.insert(cornerFromMap.end(), cornerArray[ind].begin(), cornerArray[ind].end());
        surfFromMap.insert(surfFromMap.end(), surfArray[ind].begin(), surfArray[ind].end());
    }
    last_map_corner = (int)cornerFromMap.size();
    last_map_surf = (int)surfFromMap.size();

    // LM:1814-1822
    std::vector<P4> cornerStack, surfStack;
    voxel_grid(laserCloudCornerLast, cfg.line_res, cfg.voxel_stable != 0, cornerStack);
    voxel_grid(laserCloudSurfLast, cfg.plane_res, cfg.voxel_stable != 0, surfStack);
    last_stack_corner = (int)cornerStack.size();
    last_stack_surf = (int)surfStack.size();

    int rc = 0;
    last_corner_num = last_surf_num = 0;
    if (cornerFromMap.size() > 10 && surfFromMap.size() > 50) {  // LM:1826
        KdTree kdCorner, kdSurf;  // LM:1830-1831
        kdCorner.build(cornerFromMap);
        kdSurf.build(surfFromMap);
        int pointSearchInd[5] = {0, 0, 0, 0, 0};
        float pointSearchSqDis[5] = {0, 0, 0, 0, 0};
        for (int iterCount = 0; iterCount < 2; iterCount++) {  // LM:1834
            std::vector<ResidualBlock> problem;
            int corner_num = 0, surf_num = 0;
            for (size_t i = 0; i < cornerStack.size(); i++) {  // LM:1877-1940
                const P4 pointOri = cornerStack[i];
                const P4 pointSel = associate_to_map(pointOri, parameters);
                const float qf[3] = {pointSel.x, pointSel.y, pointSel.z};
                kdCorner.knn(qf, 5, pointSearchInd, pointSearchSqDis);
                if (pointSearchSqDis[4] < 1.0) {
                    double near[5][3], center[3] = {0, 0, 0};
                    for (int j = 0; j < 5; j++) {
                        const P4& m = cornerFromMap[pointSearchInd[j]];
                        near[j][0] = m.x; near[j][1] = m.y; near[j][2] = m.z;
                        for (int a = 0; a < 3; ++a) center[a] = center[a] + near[j][a];
                    }
                    for (int a = 0; a < 3; ++a) center[a] = center[a] / 5.0;
                    double cov[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                    for (int j = 0; j < 5; j++) {
                        const double zm[3] = {near[j][0] - center[0], near[j][1] - center[1], near[j][2] - center[2]};
                        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) cov[r * 3 + c] = cov[r * 3 + c] + zm[r] * zm[c];
                    }
                    double ev[3], evec[9];
                    sym_eig3(cov, ev, evec);
                    const double dir[3] = {evec[2], evec[5], evec[8]};
                    if (ev[2] > 3 * ev[1]) {
                        const double cp[3] = {pointOri.x, pointOri.y, pointOri.z};
                        double pa[3], pb[3];
                        for (int a = 0; a < 3; ++a) { pa[a] = 0.1 * dir[a] + center[a]; pb[a] = -0.1 * dir[a] + center[a]; }
                        problem.push_back(make_edge(cp, pa, pb, 1.0));
                        corner_num++;
                    }
                }
            }
            // LM:2057-2072 (commented out in the reference; cfg.map_graph_vote turns it on): the plane correspondences are
            // collected as Corre_Match records (LM:1997-2007) and the voted ones get a SECOND LidarPlaneNormFactor block
            const bool map_vote = cfg.map_graph_vote > 0 && frameCount >= cfg.map_graph_vote - 1;
            std::vector<CorreMatch> correspondences;
            std::vector<ResidualBlock> vote_blocks;
            for (size_t i = 0; i < surfStack.size(); i++) {  // LM:1943-2055
                const P4 pointOri = surfStack[i];
                const P4 pointSel = associate_to_map(pointOri, parameters);
                const float qf[3] = {pointSel.x, pointSel.y, pointSel.z};
                kdSurf.knn(qf, 5, pointSearchInd, pointSearchSqDis);
                if (pointSearchSqDis[4] < 1.0) {
                    double matA0[15];
                    float cx = 0, cy = 0, cz = 0;   // LM:1949, 1960-1962: fp32 accumulation
                    for (int j = 0; j < 5; j++) {
                        const P4& m = surfFromMap[pointSearchInd[j]];
                        matA0[j * 3 + 0] = m.x; matA0[j * 3 + 1] = m.y; matA0[j * 3 + 2] = m.z;
                        cx += m.x; cy += m.y; cz += m.z;
                    }
                    cx /= 5; cy /= 5; cz /= 5;      // LM:1970-1972
                    double nrm[3];
                    plane_fit5(matA0, nrm);
                    const double nn = std::sqrt(nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2]);
                    const double negative_OA_dot_norm = 1 / nn;
                    // Eigen normalize(): if (squaredNorm > 0) v /= sqrt(squaredNorm)
                    if (nn * nn > 0.0) for (int a = 0; a < 3; ++a) nrm[a] = nrm[a] / nn;
                    bool planeValid = true;
                    for (int j = 0; j < 5; j++) {
                        const P4& m = surfFromMap[pointSearchInd[j]];
                        if (std::fabs(nrm[0] * m.x + nrm[1] * m.y + nrm[2] * m.z + negative_OA_dot_norm) > 0.2) { planeValid = false; break; }
                    }
                    if (planeValid) {
                        const double cp[3] = {pointOri.x, pointOri.y, pointOri.z};
                        if (map_vote) {
                            CorreMatch cor;
                            cor.index = (int)correspondences.size();
                            cor.src = surfStack[i];
                            cor.tgt = P4{cx, cy, cz, 0.f};
                            cor.score = 0; cor.s = 0;
                            correspondences.push_back(cor);
                            vote_blocks.push_back(make_plane_norm(cp, nrm, negative_OA_dot_norm));
                        }
                        problem.push_back(make_plane_norm(cp, nrm, negative_OA_dot_norm));
                        surf_num++;
                    }
                }
            }
            last_vote_selected = 0;
            if (map_vote) {   // LM:2057-2072
                std::vector<VertexVote> selected_idx;
                graph_vote_simple_mapping(correspondences, true, selected_idx);
                for (const VertexVote& v : selected_idx) problem.push_back(vote_blocks[v.index]);
                last_vote_selected = (int)selected_idx.size();
            }
            SolveSummary ss;
            solve(problem, parameters, parameters + 4, &ss, 4, true);  // LM:2079-2087
            last_solves.push_back(ss);
            last_corner_num = corner_num;
            last_surf_num = surf_num;
        }
    } else {
        rc = 1;  // "time Map corner and surf num are not enough", LM:2097-2100
    }

    {  // transformUpdate, LM:119-123
        const Quat<double> qw{q_w_curr[0], q_w_curr[1], q_w_curr[2], q_w_curr[3]};
        const Quat<double> qmw = qmul(qw, qinverse(qwo));
        const V3<double> r = rotate(qmw, two);
        q_wmap_wodom[0] = qmw.x; q_wmap_wodom[1] = qmw.y; q_wmap_wodom[2] = qmw.z; q_wmap_wodom[3] = qmw.w;
        t_wmap_wodom[0] = t_w_curr[0] - r.x; t_wmap_wodom[1] = t_w_curr[1] - r.y; t_wmap_wodom[2] = t_w_curr[2] - r.z;
    }

    // LM:2104-2152 insert the stacks into the cubes
    for (int pass = 0; pass < 2; ++pass) {
        const std::vector<P4>& stack = pass == 0 ? cornerStack : surfStack;
        std::vector<std::vector<P4>>& arr = pass == 0 ? cornerArray : surfArray;
        for (const P4& p : stack) {
            const P4 pointSel = associate_to_map(p, parameters);
            int I, J, K;
            cube_of(pointSel, cenW, cenH, cenD, I, J, K);
            if (I >= 0 && I < W && J >= 0 && J < H && K >= 0 && K < D) arr[I + W * J + W * H * K].push_back(pointSel);
        }
    }
    // LM:2155-2168 voxel DS of every valid cube
    for (int ind : validInd) {
        std::vector<P4> tmp;
        voxel_grid(cornerArray[ind], cfg.line_res, cfg.voxel_stable != 0, tmp);
        cornerArray[ind].swap(tmp);
        voxel_grid(surfArray[ind], cfg.plane_res, cfg.voxel_stable != 0, tmp);
        surfArray[ind].swap(tmp);
    }
    frameCount++;
    return rc;
}

}  // namespace orc
