// Synthetic spinning-LiDAR scan generator (host only, no CUDA, no oracle code).
//
// Produces the seeded "scene S" inputs of SURVEY.md §8(d): ground plane z = -1.73 m, an axis-aligned
// box room 120 x 80 x 15 m, 40 vertical cylinders (r = 0.15 m) and 20 boxes (2 x 2 x 3 m) at seeded
// positions kept >= 6 m away from the sensor paths, Gaussian range noise, rays ordered azimuth-major
// (all rings of one azimuth step, then the next step) as a spinning sensor emits them.  The reference
// has no data generator (its input is the ROS topic /rslidar_points, scanRegistration.cpp:453); ring
// elevation tables match the ring-id formulas at scanRegistration.cpp:142-169 so every ray falls in
// the middle of a ring bucket.
//
// Everything is counter-based (splitmix64 of seed/scan/ray), so any scan can be generated
// independently and reproducibly on any box with the same libm.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct Cyl { double x, y; };
struct Box { double x, y; };

struct Scene {
    uint64_t seed = ~0ull;
    std::vector<Cyl> cyl;
    std::vector<Box> box;
};

inline uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
inline double u01(uint64_t h) { return ((h >> 11) + 0.5) * (1.0 / 9007199254740992.0); }

inline double gauss(uint64_t key)
{
    const double u1 = u01(splitmix64(key));
    const double u2 = u01(splitmix64(key ^ 0xD1B54A32D192ED03ull));
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
}

const double kRoomX = 60.0, kRoomY = 40.0, kGround = -1.73, kCeil = 13.27;
const double kLoopRadius = 25.0;

// distance of (x,y) from both canonical sensor paths: mode 0 (1 m + 0.01 rad per scan, 64 scans
// starting at (-25,0) heading +x) and mode 1 (circle of radius 25 m around the origin).
double path_clearance(double x, double y)
{
    double best = std::fabs(std::hypot(x, y) - kLoopRadius);
    double px = -25.0, py = 0.0, yaw = 0.0;
    for (int k = 0; k < 64; ++k) {
        best = std::fmin(best, std::hypot(x - px, y - py));
        px += std::cos(yaw);
        py += std::sin(yaw);
        yaw += 0.01;
    }
    return best;
}

void build_scene(uint64_t seed, Scene& s)
{
    s.seed = seed;
    s.cyl.clear();
    s.box.clear();
    uint64_t ctr = splitmix64(seed ^ 0x5CE7E5CE7Eull);
    auto next = [&]() { ctr = splitmix64(ctr); return u01(ctr); };
    while (s.cyl.size() < 40) {
        const double x = (next() * 2 - 1) * (kRoomX - 2.0), y = (next() * 2 - 1) * (kRoomY - 2.0);
        if (path_clearance(x, y) >= 6.0) s.cyl.push_back({x, y});
    }
    while (s.box.size() < 20) {
        const double x = (next() * 2 - 1) * (kRoomX - 4.0), y = (next() * 2 - 1) * (kRoomY - 4.0);
        if (path_clearance(x, y) >= 7.5) s.box.push_back({x, y});
    }
}

Scene g_scene;

inline void slab(double o, double d, double lo, double hi, double& t0, double& t1)
{
    if (std::fabs(d) < 1e-300) {
        if (o < lo || o > hi) { t0 = 1e300; t1 = -1e300; }
        return;
    }
    double a = (lo - o) / d, b = (hi - o) / d;
    if (a > b) { double t = a; a = b; b = t; }
    if (a > t0) t0 = a;
    if (b < t1) t1 = b;
}

double raycast(const Scene& s, const double o[3], const double d[3])
{
    double best = 1e300;
    // room: the sensor is inside, so the exit distance of the room slab is the wall / floor / ceiling hit
    {
        double t0 = -1e300, t1 = 1e300;
        slab(o[0], d[0], -kRoomX, kRoomX, t0, t1);
        slab(o[1], d[1], -kRoomY, kRoomY, t0, t1);
        slab(o[2], d[2], kGround, kCeil, t0, t1);
        if (t1 > 0 && t1 < best) best = t1;
    }
    const double a = d[0] * d[0] + d[1] * d[1];
    for (const Cyl& c : s.cyl) {
        if (a < 1e-12) break;
        const double fx = o[0] - c.x, fy = o[1] - c.y;
        const double b = fx * d[0] + fy * d[1];
        const double cc = fx * fx + fy * fy - 0.15 * 0.15;
        const double disc = b * b - a * cc;
        if (disc < 0) continue;
        const double t = (-b - std::sqrt(disc)) / a;
        if (t > 0 && t < best) best = t;
    }
    for (const Box& bx : s.box) {
        double t0 = 0.0, t1 = 1e300;
        slab(o[0], d[0], bx.x - 1.0, bx.x + 1.0, t0, t1);
        slab(o[1], d[1], bx.y - 1.0, bx.y + 1.0, t0, t1);
        slab(o[2], d[2], kGround, kGround + 3.0, t0, t1);
        if (t0 <= t1 && t0 > 0 && t0 < best) best = t0;
    }
    return best;
}

}  // namespace

extern "C" {

// Elevation angle (degrees) of ring r for a 16/32/64-line sensor; inverse of the ring-id formulas at
// scanRegistration.cpp:144 (16), :153 (32, truncating -> rays sit mid-bucket at r+0.5) and :162 (64).
double ll_synth_ring_elevation_deg(int scan_line, int r, float lower_bound, float up_bound)
{
    if (scan_line == 16) return -15.0 + 2.0 * r;
    if (scan_line == 32) return -92.0 / 3.0 + (r + 0.5) * 4.0 / 3.0;
    return (double)lower_bound + r * ((double)up_bound - (double)lower_bound) / (scan_line - 1);
}

// Canonical sensor pose of scan k. mode 0: 1 m forward + 0.01 rad yaw per scan from (-25,0,0);
// mode 1: circle of radius 25 m, 1 m of arc per scan (yaw rate 0.04 rad / scan), plus a seeded smooth
// yaw wobble. out = {x, y, z, yaw}.
void ll_synth_pose(int mode, uint64_t seed, int k, double out[4])
{
    if (mode == 0) {
        double px = -25.0, py = 0.0, yaw = 0.0;
        for (int i = 0; i < k; ++i) { px += std::cos(yaw); py += std::sin(yaw); yaw += 0.01; }
        out[0] = px; out[1] = py; out[2] = 0.0; out[3] = yaw;
    } else {
        const double th = (double)k / kLoopRadius;
        const double wob = 0.01 * std::sin(0.37 * k + 6.283185307179586 * u01(splitmix64(seed ^ 0xA5A5ull)));
        out[0] = kLoopRadius * std::cos(th);
        out[1] = kLoopRadius * std::sin(th);
        out[2] = 0.0;
        out[3] = th + 1.5707963267948966 + wob;
    }
}

// Generates one scan into out (float4 x,y,z,0 per point, sensor frame). Returns the number of points
// written (<= cap), or -1 on bad arguments. Points closer than 0.5 m or farther than 120 m are dropped.
int ll_synth_scan(int scan_line, int az_steps, uint64_t seed, uint64_t scan_id, const double pose[4],
                  float lower_bound, float up_bound, float noise_sigma, float* out, int cap)
{
    if ((scan_line != 16 && scan_line != 32 && scan_line != 64) || az_steps <= 0 || !out) return -1;
    if (g_scene.seed != seed) build_scene(seed, g_scene);
    const double cy = std::cos(pose[3]), sy = std::sin(pose[3]);
    const double o[3] = {pose[0], pose[1], pose[2]};
    std::vector<double> ce(scan_line), se(scan_line);
    for (int r = 0; r < scan_line; ++r) {
        const double e = ll_synth_ring_elevation_deg(scan_line, r, lower_bound, up_bound) * 0.017453292519943295;
        ce[r] = std::cos(e);
        se[r] = std::sin(e);
    }
    int n = 0;
    for (int k = 0; k < az_steps; ++k) {
        const double phi = -6.283185307179586 * (double)k / (double)az_steps;  // clockwise spin
        const double cphi = std::cos(phi), sphi = std::sin(phi);
        for (int r = 0; r < scan_line; ++r) {
            const double ds[3] = {ce[r] * cphi, ce[r] * sphi, se[r]};
            const double dw[3] = {cy * ds[0] - sy * ds[1], sy * ds[0] + cy * ds[1], ds[2]};
            double t = raycast(g_scene, o, dw);
            if (!(t < 1e200)) continue;
            const uint64_t key = splitmix64(seed ^ (scan_id * 0x9E3779B97F4A7C15ull)) ^
                                 ((uint64_t)(k * scan_line + r) * 0xC2B2AE3D27D4EB4Full);
            t += (double)noise_sigma * gauss(key);
            if (t < 0.5 || t > 120.0) continue;
            if (n >= cap) return n;
            out[4 * n + 0] = (float)(t * ds[0]);
            out[4 * n + 1] = (float)(t * ds[1]);
            out[4 * n + 2] = (float)(t * ds[2]);
            out[4 * n + 3] = 0.0f;
            ++n;
        }
    }
    return n;
}

}  // extern "C"
