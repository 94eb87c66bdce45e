// ORACLE (test infrastructure) — forward-mode dual numbers and the Eigen quaternion/vector snippets the
// reference's cost functors use (lidarFactor.hpp), restated so the functors can be written once for
// T = double (cost-only evaluation) and T = Jet<7> (ceres::AutoDiffCostFunction<.., 4, 3>).
//
// Restates: ceres/jet.h (Jet arithmetic, sqrt, sin, acos, abs, comparisons on the scalar part) and
// Eigen 3.3 Geometry/Quaternion.h (`_transformVector`, `slerp`, product, inverse) — SURVEY.md A.3/A.4.
#pragma once
#include <cmath>
#include <limits>

namespace orc {

template <int N>
struct Jet {
    double a;
    double v[N];
    Jet() : a(0.0) { for (int i = 0; i < N; ++i) v[i] = 0.0; }
    Jet(double s) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; }  // NOLINT (implicit like ceres)
    Jet(double s, int k) : a(s) { for (int i = 0; i < N; ++i) v[i] = 0.0; v[k] = 1.0; }
};

template <int N> inline Jet<N> operator+(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a + g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] + g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a - g.a; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] - g.v[i]; return h; }
template <int N> inline Jet<N> operator-(const Jet<N>& f) { Jet<N> h; h.a = -f.a; for (int i = 0; i < N; ++i) h.v[i] = -f.v[i]; return h; }
template <int N> inline Jet<N> operator*(const Jet<N>& f, const Jet<N>& g) { Jet<N> h; h.a = f.a * g.a; for (int i = 0; i < N; ++i) h.v[i] = f.a * g.v[i] + f.v[i] * g.a; return h; }
template <int N> inline Jet<N> operator/(const Jet<N>& f, const Jet<N>& g)
{   // ceres/jet.h: g_a_inverse, f_a_by_g_a, (f.v - f_a_by_g_a * g.v) * g_a_inverse
    Jet<N> h;
    const double g_a_inverse = 1.0 / g.a;
    const double f_a_by_g_a = f.a * g_a_inverse;
    h.a = f_a_by_g_a;
    for (int i = 0; i < N; ++i) h.v[i] = (f.v[i] - f_a_by_g_a * g.v[i]) * g_a_inverse;
    return h;
}
template <int N> inline Jet<N> operator*(const Jet<N>& f, double s) { Jet<N> h; h.a = f.a * s; for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * s; return h; }
template <int N> inline Jet<N> operator*(double s, const Jet<N>& f) { return f * s; }
template <int N> inline Jet<N> operator+(const Jet<N>& f, double s) { Jet<N> h = f; h.a += s; return h; }
template <int N> inline Jet<N> operator-(double s, const Jet<N>& f) { Jet<N> h = -f; h.a += s; return h; }
template <int N> inline bool operator<(const Jet<N>& f, const Jet<N>& g) { return f.a < g.a; }
template <int N> inline bool operator>=(const Jet<N>& f, const Jet<N>& g) { return f.a >= g.a; }

template <int N> inline Jet<N> jsqrt(const Jet<N>& f) { Jet<N> h; const double t = std::sqrt(f.a); h.a = t; const double two_a_inverse = 1.0 / (2.0 * t); for (int i = 0; i < N; ++i) h.v[i] = f.v[i] * two_a_inverse; return h; }
template <int N> inline Jet<N> jsin(const Jet<N>& f) { Jet<N> h; h.a = std::sin(f.a); const double c = std::cos(f.a); for (int i = 0; i < N; ++i) h.v[i] = c * f.v[i]; return h; }
template <int N> inline Jet<N> jacos(const Jet<N>& f) { Jet<N> h; h.a = std::acos(f.a); const double t = -1.0 / std::sqrt(1.0 - f.a * f.a); for (int i = 0; i < N; ++i) h.v[i] = t * f.v[i]; return h; }
template <int N> inline Jet<N> jabs(const Jet<N>& f) { return f.a < 0.0 ? -f : f; }
inline double jsqrt(double x) { return std::sqrt(x); }
inline double jsin(double x) { return std::sin(x); }
inline double jacos(double x) { return std::acos(x); }
inline double jabs(double x) { return std::fabs(x); }

template <typename T> struct V3 { T x, y, z; };
template <typename T> inline V3<T> operator+(const V3<T>& a, const V3<T>& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
template <typename T> inline V3<T> operator-(const V3<T>& a, const V3<T>& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
template <typename T> inline V3<T> cross(const V3<T>& a, const V3<T>& b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
template <typename T> inline T dot(const V3<T>& a, const V3<T>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> inline T norm(const V3<T>& a) { return jsqrt(dot(a, a)); }

// Eigen::Quaternion<T>, coefficients stored x,y,z,w; constructor order in the reference is (w,x,y,z).
template <typename T> struct Quat { T x, y, z, w; };

// QuaternionBase::_transformVector (Eigen 3.3): uv = 2 (u x v); v + w*uv + u x uv.
template <typename T> inline V3<T> rotate(const Quat<T>& q, const V3<T>& v)
{
    const V3<T> u{q.x, q.y, q.z};
    V3<T> uv = cross(u, v);
    uv = uv + uv;
    const V3<T> wuv{q.w * uv.x, q.w * uv.y, q.w * uv.z};
    return (v + wuv) + cross(u, uv);
}

// QuaternionBase::slerp(t, other) with *this = identity (lidarFactor.hpp:25-26, laserOdometry.cpp:86).
template <typename T> inline Quat<T> identity_slerp(const T& t, const Quat<T>& other)
{
    const T one = T(1.0) - T(std::numeric_limits<double>::epsilon());
    const T d = T(0.0) * other.x + T(0.0) * other.y + T(0.0) * other.z + T(1.0) * other.w;  // this->dot(other)
    const T absD = jabs(d);
    T scale0, scale1;
    if (absD >= one) {
        scale0 = T(1.0) - t;
        scale1 = t;
    } else {
        const T theta = jacos(absD);
        const T sinTheta = jsin(theta);
        scale0 = jsin((T(1.0) - t) * theta) / sinTheta;
        scale1 = jsin((t * theta)) / sinTheta;
    }
    if (d < T(0.0)) scale1 = -scale1;
    // scale0 * coeffs() + scale1 * other.coeffs(), identity coeffs = (0,0,0,1)
    return {scale0 * T(0.0) + scale1 * other.x, scale0 * T(0.0) + scale1 * other.y, scale0 * T(0.0) + scale1 * other.z,
            scale0 * T(1.0) + scale1 * other.w};
}

// Eigen quat_product (Hamilton), and inverse = conjugate / squaredNorm.
inline Quat<double> qmul(const Quat<double>& a, const Quat<double>& b)
{
    return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
            a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
inline Quat<double> qinverse(const Quat<double>& q)
{
    const double n2 = q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    if (n2 > 0.0) return {-q.x / n2, -q.y / n2, -q.z / n2, q.w / n2};
    return {0.0, 0.0, 0.0, 0.0};
}

}  // namespace orc
