// ORACLE (test infrastructure) — scanRegistration.cpp:87-428 `laserCloudHandler` restated.
//
// fp conventions (SURVEY.md §3.1 notes): `atan`/`sqrt` at SR:139 resolve to the float overloads
// (ROS/PCL headers pull <math.h>'s C++ overloads into the global namespace), `-atan2(y,x)` is
// std::atan2(float,float); comparisons against M_PI promote to double exactly as written in the
// reference.  Build with -ffp-contract=off so fp32 sums are never fused.
#include "orc_api.h"

#include <algorithm>
#include <cmath>

namespace orc {

int extract_features(const float* pts, int n_in, int stride_floats, const Config& cfg, Features& out)
{
    const int N_SCANS = cfg.scan_line;
    if (N_SCANS != 16 && N_SCANS != 32 && N_SCANS != 64) return -1;  // SR:447-451
    out = Features();

    // SR:105-110  fromROSMsg -> PointXYZ, removeNaNFromPointCloud, removeClosedPointCloud(thres)
    struct P3 { float x, y, z; };
    std::vector<P3> in;
    in.reserve(n_in);
    const float thres = cfg.minimum_range;
    for (int i = 0; i < n_in; ++i) {
        const float x = pts[(size_t)i * stride_floats + 0], y = pts[(size_t)i * stride_floats + 1],
                    z = pts[(size_t)i * stride_floats + 2];
        if (!std::isfinite(x) || !std::isfinite(y) || !std::isfinite(z)) continue;
        if (x * x + y * y + z * z < thres * thres) continue;  // SR:72
        in.push_back({x, y, z});
    }
    int cloudSize = (int)in.size();
    if (cloudSize == 0) return -2;  // reference would index points[0] (UB); we report it

    // SR:114-126
    float startOri = -std::atan2(in[0].y, in[0].x);
    float endOri = -std::atan2(in[cloudSize - 1].y, in[cloudSize - 1].x) + 2 * M_PI;
    if (endOri - startOri > 3 * M_PI)
        endOri -= 2 * M_PI;
    else if (endOri - startOri < M_PI)
        endOri += 2 * M_PI;

    const float lowerBound = cfg.lower_bound, upBound = cfg.up_bound;
    const float _factor = (N_SCANS - 1) / (upBound - lowerBound);  // SR:441
    const double scanPeriod = 0.1;                                  // SR:28

    bool halfPassed = false;
    int count = cloudSize;
    std::vector<std::vector<P4>> laserCloudScans(N_SCANS);
    for (int i = 0; i < cloudSize; i++) {  // SR:133-210
        P4 point;
        point.x = in[i].x;
        point.y = in[i].y;
        point.z = in[i].z;
        float angle = std::atan(point.z / std::sqrt(point.x * point.x + point.y * point.y)) * 180 / M_PI;
        int scanID = 0;
        if (N_SCANS == 16) {
            scanID = int((angle + 15) / 2 + 0.5);
            if (scanID > (N_SCANS - 1) || scanID < 0) { count--; continue; }
        } else if (N_SCANS == 32) {
            scanID = int((angle + 92.0 / 3.0) * 3.0 / 4.0);
            if (scanID > (N_SCANS - 1) || scanID < 0) { count--; continue; }
        } else {
            scanID = int((angle - lowerBound) * _factor + 0.5);
            if (scanID >= N_SCANS || scanID < 0) { count--; continue; }
        }

        float ori = -std::atan2(point.y, point.x);
        if (!halfPassed) {
            if (ori < startOri - M_PI / 2)
                ori += 2 * M_PI;
            else if (ori > startOri + M_PI * 3 / 2)
                ori -= 2 * M_PI;
            if (ori - startOri > M_PI) halfPassed = true;
        } else {
            ori += 2 * M_PI;
            if (ori < endOri - M_PI * 3 / 2)
                ori += 2 * M_PI;
            else if (ori > endOri + M_PI / 2)
                ori -= 2 * M_PI;
        }
        float relTime = (ori - startOri) / (endOri - startOri);
        point.i = scanID + scanPeriod * relTime;
        laserCloudScans[scanID].push_back(point);
    }
    cloudSize = count;

    // SR:215-221
    std::vector<int> scanStartInd(N_SCANS, 0), scanEndInd(N_SCANS, 0);
    std::vector<P4>& laserCloud = out.full;
    laserCloud.reserve(cloudSize);
    out.ring_begin.assign(N_SCANS + 1, 0);
    for (int i = 0; i < N_SCANS; i++) {
        out.ring_begin[i] = (int)laserCloud.size();
        scanStartInd[i] = (int)laserCloud.size() + 5;
        laserCloud.insert(laserCloud.end(), laserCloudScans[i].begin(), laserCloudScans[i].end());
        scanEndInd[i] = (int)laserCloud.size() - 6;
    }
    out.ring_begin[N_SCANS] = (int)laserCloud.size();

    // SR:225-235
    std::vector<float>& cloudCurvature = out.curvature;
    std::vector<int>& cloudLabel = out.label;
    cloudCurvature.assign(cloudSize, 0.f);
    cloudLabel.assign(cloudSize, 0);
    std::vector<int> cloudSortInd(cloudSize, 0), cloudNeighborPicked(cloudSize, 0);
    // 11-tap stencil in the reference's exact left-to-right fp32 order (SR:228-230):
    // ((((p[-5]+p[-4])+p[-3])+p[-2])+p[-1]) - 10*p[0], then + p[+1] ... + p[+5].
    auto stencil = [&](int i, float P4::*ax) {
        float acc = laserCloud[i - 5].*ax;
        for (int k = -4; k <= -1; ++k) acc = acc + laserCloud[i + k].*ax;
        acc = acc - 10 * (laserCloud[i].*ax);
        for (int k = 1; k <= 5; ++k) acc = acc + laserCloud[i + k].*ax;
        return acc;
    };
    for (int i = 5; i < cloudSize - 5; i++) {
        const float diffX = stencil(i, &P4::x), diffY = stencil(i, &P4::y), diffZ = stencil(i, &P4::z);
        cloudCurvature[i] = diffX * diffX + diffY * diffY + diffZ * diffZ;
        cloudSortInd[i] = i;
        cloudNeighborPicked[i] = 0;
        cloudLabel[i] = 0;
    }

    auto comp = [&](int i, int j) { return cloudCurvature[i] < cloudCurvature[j]; };  // SR:42
    auto gap2 = [&](int a, int b) {  // SR:290-293: squared distance of consecutive points, fp32
        float diffX = laserCloud[a].x - laserCloud[b].x;
        float diffY = laserCloud[a].y - laserCloud[b].y;
        float diffZ = laserCloud[a].z - laserCloud[b].z;
        return diffX * diffX + diffY * diffY + diffZ * diffZ;
    };

    out.less_flat_ring_count.assign(N_SCANS, 0);
    for (int i = 0; i < N_SCANS; i++) {  // SR:246-377
        if (scanEndInd[i] - scanStartInd[i] < 6) continue;
        std::vector<P4> surfPointsLessFlatScan;
        for (int j = 0; j < 6; j++) {
            int sp = scanStartInd[i] + (scanEndInd[i] - scanStartInd[i]) * j / 6;
            int ep = scanStartInd[i] + (scanEndInd[i] - scanStartInd[i]) * (j + 1) / 6 - 1;

            std::sort(cloudSortInd.begin() + sp, cloudSortInd.begin() + ep + 1, comp);  // SR:257
            for (int k = sp; k < ep; ++k)
                if (cloudCurvature[cloudSortInd[k]] == cloudCurvature[cloudSortInd[k + 1]]) out.sort_ties++;

            int largestPickedNum = 0;
            for (int k = ep; k >= sp; k--) {  // SR:261-313
                int ind = cloudSortInd[k];
                if (cloudNeighborPicked[ind] == 0 && cloudCurvature[ind] > 0.1) {
                    largestPickedNum++;
                    if (largestPickedNum <= 2) {
                        cloudLabel[ind] = 2;
                        out.sharp_idx.push_back(ind);
                        out.less_sharp_idx.push_back(ind);
                    } else if (largestPickedNum <= 20) {
                        cloudLabel[ind] = 1;
                        out.less_sharp_idx.push_back(ind);
                    } else {
                        break;
                    }
                    cloudNeighborPicked[ind] = 1;
                    for (int l = 1; l <= 5; l++) {
                        if (gap2(ind + l, ind + l - 1) > 0.05) break;
                        cloudNeighborPicked[ind + l] = 1;
                    }
                    for (int l = -1; l >= -5; l--) {
                        if (gap2(ind + l, ind + l + 1) > 0.05) break;
                        cloudNeighborPicked[ind + l] = 1;
                    }
                }
            }

            int smallestPickedNum = 0;
            for (int k = sp; k <= ep; k++) {  // SR:316-359
                int ind = cloudSortInd[k];
                if (cloudNeighborPicked[ind] == 0 && cloudCurvature[ind] < 0.1) {
                    cloudLabel[ind] = -1;
                    out.flat_idx.push_back(ind);
                    smallestPickedNum++;
                    if (smallestPickedNum >= 4) break;
                    cloudNeighborPicked[ind] = 1;
                    for (int l = 1; l <= 5; l++) {
                        if (gap2(ind + l, ind + l - 1) > 0.05) break;
                        cloudNeighborPicked[ind + l] = 1;
                    }
                    for (int l = -1; l >= -5; l--) {
                        if (gap2(ind + l, ind + l + 1) > 0.05) break;
                        cloudNeighborPicked[ind + l] = 1;
                    }
                }
            }

            for (int k = sp; k <= ep; k++)  // SR:361-367
                if (cloudLabel[k] <= 0) surfPointsLessFlatScan.push_back(laserCloud[k]);
        }
        std::vector<P4> ds;  // SR:370-376
        voxel_grid(surfPointsLessFlatScan, 0.2f, cfg.voxel_stable != 0, ds);
        out.less_flat_ring_count[i] = (int)ds.size();
        out.less_flat.insert(out.less_flat.end(), ds.begin(), ds.end());
    }

    for (int ind : out.sharp_idx) out.sharp.push_back(laserCloud[ind]);
    for (int ind : out.less_sharp_idx) out.less_sharp.push_back(laserCloud[ind]);
    for (int ind : out.flat_idx) out.flat.push_back(laserCloud[ind]);
    return 0;
}

}  // namespace orc
