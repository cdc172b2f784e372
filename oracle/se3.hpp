// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into, imported by, or executed from the product path
// (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it).
//
// SE(3)/SO(3) arithmetic restated from the published Sophus 1.22.11 / Eigen 3.4 algorithms, which the
// reference pulls in as un-vendored third-party deps (cpp/sage_icp/3rdparty/sophus/sophus.cmake:29,
// 3rdparty/eigen/eigen.cmake:35). Reference call sites this file stands in for:
//   SE3d::exp                 core/Registration.cpp:93, core/Deskew.cpp:43
//   SE3d::log                 core/Registration.cpp:137, core/Deskew.cpp:40
//   SE3d * SE3d, inverse()    pipeline/sageICP.cpp:76,90,114,119; core/Registration.cpp:135,140
//   SE3d * Vector3d           core/Registration.cpp:106, core/VoxelHashMap.cpp:154
//   SO3d::hat                 core/Registration.cpp:68
//   AngleAxisd(R).angle()     core/Threshold.cpp:30
//   Matrix6d::ldlt().solve    core/Registration.cpp:92
// PARITY UNPINNED: the reference has no tests/golden vectors; these formulas are
// cross-checked against scipy (expm/logm/Rotation) in tests/test_oracle_se3.py instead.
#pragma once
#include <cmath>
#include <cstring>

namespace orc {

constexpr double kSophusEps = 1e-10;  // Sophus::Constants<double>::epsilon()

struct Vec3 {
    double x, y, z;
};
inline Vec3 operator+(const Vec3 &a, const Vec3 &b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline Vec3 operator-(const Vec3 &a, const Vec3 &b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline Vec3 operator*(double s, const Vec3 &a) { return {s * a.x, s * a.y, s * a.z}; }
inline Vec3 cross(const Vec3 &a, const Vec3 &b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
// Eigen fixed-size-3 squaredNorm reduction order: (x*x + y*y) + z*z
inline double sqnorm(const Vec3 &a) { return (a.x * a.x + a.y * a.y) + a.z * a.z; }
inline double norm(const Vec3 &a) { return std::sqrt(sqnorm(a)); }

struct Mat3 {
    double m[3][3];
};
inline Mat3 mat3_identity() { return {{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}}; }
inline Mat3 mat3_mul(const Mat3 &a, const Mat3 &b) {
    Mat3 c;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) c.m[i][j] = a.m[i][0] * b.m[0][j] + a.m[i][1] * b.m[1][j] + a.m[i][2] * b.m[2][j];
    return c;
}
inline Vec3 mat3_vec(const Mat3 &a, const Vec3 &v) {
    return {a.m[0][0] * v.x + a.m[0][1] * v.y + a.m[0][2] * v.z,
            a.m[1][0] * v.x + a.m[1][1] * v.y + a.m[1][2] * v.z,
            a.m[2][0] * v.x + a.m[2][1] * v.y + a.m[2][2] * v.z};
}
// SO3d::hat (core/Registration.cpp:68)
inline Mat3 hat(const Vec3 &w) { return {{{0, -w.z, w.y}, {w.z, 0, -w.x}, {-w.y, w.x, 0}}}; }

struct Quat {
    double w = 1, x = 0, y = 0, z = 0;
};

inline Quat quat_normalized(Quat q) {
    const double n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
    return {q.w / n, q.x / n, q.y / n, q.z / n};
}

// Hamilton product followed by renormalisation (Sophus SO3 product builds a new SO3 from the raw
// quaternion product, whose constructor normalises).
inline Quat quat_mul(const Quat &a, const Quat &b) {
    Quat r{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
           a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
    return quat_normalized(r);
}
inline Quat quat_conj(const Quat &q) { return {q.w, -q.x, -q.y, -q.z}; }

// Eigen QuaternionBase::_transformVector: uv = 2 * (q.vec x v); v + w*uv + q.vec x uv
inline Vec3 quat_rotate(const Quat &q, const Vec3 &v) {
    const Vec3 qv{q.x, q.y, q.z};
    Vec3 uv = cross(qv, v);
    uv = uv + uv;
    return (v + q.w * uv) + cross(qv, uv);
}

// Eigen QuaternionBase::toRotationMatrix
inline Mat3 quat_matrix(const Quat &q) {
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
    const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    return {{{1 - (tyy + tzz), txy - twz, txz + twy}, {txy + twz, 1 - (txx + tzz), tyz - twx}, {txz - twy, tyz + twx, 1 - (txx + tyy)}}};
}

// Eigen quaternion-from-rotation-matrix (Shepperd / "Ken Shoemake" branch form used by Eigen)
inline Quat quat_from_matrix(const Mat3 &a) {
    Quat q;
    double t = a.m[0][0] + a.m[1][1] + a.m[2][2];
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (a.m[2][1] - a.m[1][2]) * t;
        q.y = (a.m[0][2] - a.m[2][0]) * t;
        q.z = (a.m[1][0] - a.m[0][1]) * t;
    } else {
        int i = 0;
        if (a.m[1][1] > a.m[0][0]) i = 1;
        if (a.m[2][2] > a.m[i][i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(a.m[i][i] - a.m[j][j] - a.m[k][k] + 1.0);
        double v[3];
        v[i] = 0.5 * t;
        t = 0.5 / t;
        q.w = (a.m[k][j] - a.m[j][k]) * t;
        v[j] = (a.m[j][i] + a.m[i][j]) * t;
        v[k] = (a.m[k][i] + a.m[i][k]) * t;
        q.x = v[0], q.y = v[1], q.z = v[2];
    }
    return q;
}

// Eigen::AngleAxisd(Matrix3d).angle(): via quaternion, angle = 2*atan2(|vec|, |w|) in [0, pi]
inline double angle_of_rotation_matrix(const Mat3 &R) {
    const Quat q = quat_from_matrix(R);
    const double n = std::sqrt((q.x * q.x + q.y * q.y) + q.z * q.z);
    if (n != 0.0) return 2.0 * std::atan2(n, std::fabs(q.w));
    return 0.0;
}

struct SE3 {
    Quat q;
    Vec3 t{0, 0, 0};
};

inline SE3 se3_mul(const SE3 &a, const SE3 &b) { return {quat_mul(a.q, b.q), a.t + quat_rotate(a.q, b.t)}; }
inline SE3 se3_inverse(const SE3 &a) {
    const Quat qi = quat_conj(a.q);
    return {qi, quat_rotate(qi, -1.0 * a.t)};
}
inline Vec3 se3_act(const SE3 &a, const Vec3 &p) { return quat_rotate(a.q, p) + a.t; }

// Sophus SO3::expAndTheta
inline Quat so3_exp(const Vec3 &omega, double *theta_out) {
    const double theta_sq = sqnorm(omega);
    double imag, real, theta;
    if (theta_sq < kSophusEps * kSophusEps) {
        theta = 0;
        const double theta_po4 = theta_sq * theta_sq;
        imag = 0.5 - (1.0 / 48.0) * theta_sq + (1.0 / 3840.0) * theta_po4;
        real = 1.0 - (1.0 / 8.0) * theta_sq + (1.0 / 384.0) * theta_po4;
    } else {
        theta = std::sqrt(theta_sq);
        const double half = 0.5 * theta;
        imag = std::sin(half) / theta;
        real = std::cos(half);
    }
    if (theta_out) *theta_out = theta;
    // SO3 is built from the raw quaternion; the constructor normalises
    return quat_normalized({real, imag * omega.x, imag * omega.y, imag * omega.z});
}

// Sophus SE3::exp; tangent = (upsilon, omega)
inline SE3 se3_exp(const double xi[6]) {
    const Vec3 ups{xi[0], xi[1], xi[2]}, omega{xi[3], xi[4], xi[5]};
    double theta;
    const Quat q = so3_exp(omega, &theta);
    const Mat3 Om = hat(omega), Om2 = mat3_mul(Om, Om);
    Mat3 V;
    if (theta < kSophusEps) {
        V = quat_matrix(q);
    } else {
        const double th2 = theta * theta;
        const double a = (1.0 - std::cos(theta)) / th2, b = (theta - std::sin(theta)) / (th2 * theta);
        const Mat3 I = mat3_identity();
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) V.m[i][j] = I.m[i][j] + a * Om.m[i][j] + b * Om2.m[i][j];
    }
    return {q, mat3_vec(V, ups)};
}

// Sophus SO3::logAndTheta
inline Vec3 so3_log(const Quat &q, double *theta_out) {
    const double sq_n = (q.x * q.x + q.y * q.y) + q.z * q.z;
    const double w = q.w;
    double two_atan, theta;
    if (sq_n < kSophusEps * kSophusEps) {
        const double sq_w = w * w;
        two_atan = 2.0 / w - (2.0 / 3.0) * sq_n / (w * sq_w);
        theta = 2.0 * sq_n / w;
    } else {
        const double n = std::sqrt(sq_n);
        const double at = (w < 0) ? std::atan2(-n, -w) : std::atan2(n, w);
        two_atan = 2.0 * at / n;
        theta = two_atan * n;
    }
    if (theta_out) *theta_out = theta;
    return {two_atan * q.x, two_atan * q.y, two_atan * q.z};
}

// Sophus SE3::log -> (upsilon, omega)
inline void se3_log(const SE3 &T, double xi[6]) {
    double theta;
    const Vec3 omega = so3_log(T.q, &theta);
    const Mat3 Om = hat(omega), Om2 = mat3_mul(Om, Om), I = mat3_identity();
    Mat3 Vinv;
    double c;
    if (std::fabs(theta) < kSophusEps) {
        c = 1.0 / 12.0;
    } else {
        const double half = 0.5 * theta;
        c = (1.0 - theta * std::cos(half) / (2.0 * std::sin(half))) / (theta * theta);
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) Vinv.m[i][j] = I.m[i][j] - 0.5 * Om.m[i][j] + c * Om2.m[i][j];
    const Vec3 u = mat3_vec(Vinv, T.t);
    xi[0] = u.x, xi[1] = u.y, xi[2] = u.z, xi[3] = omega.x, xi[4] = omega.y, xi[5] = omega.z;
}

// Eigen LDLT (symmetric, diagonal pivoting on the largest |A_ii|) solve of a 6x6 system A x = b.
// Restated algorithm; agrees with any f64 Cholesky to ~1e-12 relative on SPD input.
inline void ldlt6_solve(const double Ain[6][6], const double bin[6], double x[6]) {
    double A[6][6], b[6];
    int perm[6];
    std::memcpy(A, Ain, sizeof(A));
    for (int i = 0; i < 6; ++i) perm[i] = i, b[i] = bin[i];
    for (int k = 0; k < 6; ++k) {
        int p = k;
        double best = std::fabs(A[k][k]);
        for (int i = k + 1; i < 6; ++i)
            if (std::fabs(A[i][i]) > best) best = std::fabs(A[i][i]), p = i;
        if (p != k) {  // symmetric row/col swap
            for (int j = 0; j < 6; ++j) std::swap(A[k][j], A[p][j]);
            for (int i = 0; i < 6; ++i) std::swap(A[i][k], A[i][p]);
            std::swap(perm[k], perm[p]);
        }
        const double d = A[k][k];
        if (d == 0.0) continue;
        for (int i = k + 1; i < 6; ++i) A[i][k] /= d;  // L column
        for (int i = k + 1; i < 6; ++i)
            for (int j = k + 1; j <= i; ++j) {
                A[i][j] -= A[i][k] * d * A[j][k];
                A[j][i] = A[i][j];
            }
    }
    double y[6];
    for (int i = 0; i < 6; ++i) y[i] = b[perm[i]];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < i; ++j) y[i] -= A[i][j] * y[j];
    for (int i = 0; i < 6; ++i) y[i] = (A[i][i] != 0.0) ? y[i] / A[i][i] : 0.0;
    for (int i = 5; i >= 0; --i)
        for (int j = i + 1; j < 6; ++j) y[i] -= A[j][i] * y[j];
    for (int i = 0; i < 6; ++i) x[perm[i]] = y[i];
}

}  // namespace orc
