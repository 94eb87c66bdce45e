// ll_run — ROS-free C++ host driver over the C ABI (include/lightloam_b200.h).
//
// Plays a sequence of scans (KITTI-style .bin files of float32 x,y,z,i records — the format kittiHelper.cpp:22-32 reads —
// or the seeded synthetic generator) through the fused pipeline and writes the trajectory in the reference's
// result-file format: per scan the top 3 x 4 of H_init^-1 * H as 12 numbers, scientific, precision 6
// (laserMapping.cpp:2284-2325).  This is the C++ side a maintainer starts from; the ROS nodes are in nodes/.
//
//   ll_run --lines 64 --scans 50 [--mapping] [--out traj.txt] [--bin-dir DIR]
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/lightloam_b200.h"

extern "C" {
int ll_synth_scan(int scan_line, int az_steps, uint64_t seed, uint64_t scan_id, const double pose[4], float lower_bound, float up_bound,
                  float noise_sigma, float* out, int cap);
void ll_synth_pose(int mode, uint64_t seed, int k, double out[4]);
}

static void quat_to_rot(const double q[4], double R[9])
{   // Eigen toRotationMatrix(), q = x,y,z,w
    const double x = q[0], y = q[1], z = q[2], w = q[3];
    R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w); R[2] = 2 * (x * z + y * w);
    R[3] = 2 * (x * y + z * w); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
    R[6] = 2 * (x * z - y * w); R[7] = 2 * (y * z + x * w); R[8] = 1 - 2 * (x * x + y * y);
}

int main(int argc, char** argv)
{
    int lines = 64, scans = 20, mapping = 0;
    std::string out_path, bin_dir;
    for (int i = 1; i < argc; ++i) {
        if (!strcmp(argv[i], "--lines") && i + 1 < argc) lines = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--scans") && i + 1 < argc) scans = atoi(argv[++i]);
        else if (!strcmp(argv[i], "--mapping")) mapping = 1;
        else if (!strcmp(argv[i], "--out") && i + 1 < argc) out_path = argv[++i];
        else if (!strcmp(argv[i], "--bin-dir") && i + 1 < argc) bin_dir = argv[++i];
    }
    ll_config cfg;
    ll_default_config(&cfg, lines);
    cfg.enable_mapping = mapping;
    cfg.map_capacity = 1 << 20;
    ll_ctx* ctx = nullptr;
    int rc = ll_create(&cfg, &ctx);
    if (rc) { fprintf(stderr, "ll_create: %s\n", ll_strerror(rc)); return 1; }
    const int az = lines == 64 ? 2031 : (lines == 32 ? 2170 : 1000);
    std::vector<float> buf((size_t)cfg.max_points * 4);
    FILE* fo = out_path.empty() ? stdout : fopen(out_path.c_str(), "w");
    float Hinit[16];
    bool init = true;
    for (int k = 0; k < scans; ++k) {
        int n = 0;
        if (!bin_dir.empty()) {
            char path[1024];
            snprintf(path, sizeof(path), "%s/%06d.bin", bin_dir.c_str(), k);
            FILE* f = fopen(path, "rb");
            if (!f) break;
            n = (int)(fread(buf.data(), 16, cfg.max_points, f));
            fclose(f);
        } else {
            double pose[4];
            ll_synth_pose(0, 20240919, k, pose);
            n = ll_synth_scan(lines, az, 20240919, (uint64_t)k, pose, cfg.lower_bound, cfg.up_bound, 0.02f, buf.data(), cfg.max_points);
        }
        ll_cloud_view v{buf.data(), n, 16};
        double poses[14];
        rc = ll_process_scans(ctx, 1, &v, poses);
        if (rc < 0) { fprintf(stderr, "scan %d: %s (%s)\n", k, ll_strerror(rc), ll_last_error(ctx)); return 2; }
        const double* q = poses + 7;
        const double* t = poses + 11;
        double R[9];
        quat_to_rot(q, R);
        float H[16] = {(float)R[0], (float)R[1], (float)R[2], (float)t[0], (float)R[3], (float)R[4], (float)R[5], (float)t[1],
                       (float)R[6], (float)R[7], (float)R[8], (float)t[2], 0, 0, 0, 1};
        if (init) { memcpy(Hinit, H, sizeof(H)); init = false; }
        // H_init^-1 * H for a rigid transform: [R0^T | -R0^T t0]
        float Ri[9], ti[3];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) Ri[r * 3 + c] = Hinit[c * 4 + r];
        for (int r = 0; r < 3; ++r) ti[r] = -(Ri[r * 3] * Hinit[3] + Ri[r * 3 + 1] * Hinit[7] + Ri[r * 3 + 2] * Hinit[11]);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 4; ++c) {
                float v2 = Ri[r * 3] * H[c] + Ri[r * 3 + 1] * H[4 + c] + Ri[r * 3 + 2] * H[8 + c] + (c == 3 ? ti[r] : 0.f);
                fprintf(fo, (r == 2 && c == 3) ? "%.6e\n" : "%.6e ", v2);
            }
    }
    ll_stats st;
    ll_get_last_stats(ctx, &st);
    fprintf(stderr, "frames %d, last scan: %d pts, %d sharp, %d flat, %d launches\n", st.frame, st.n_full, st.n_sharp, st.n_flat, st.kernel_launches);
    if (fo != stdout) fclose(fo);
    ll_destroy(ctx);
    return 0;
}
