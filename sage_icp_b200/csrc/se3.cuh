// SE(3) arithmetic for host and device (f64).  Product code: independent of oracle/.
// Follows the published Sophus 1.22.11 / Eigen 3.4 formulas the reference relies on:
//   SE3d::exp / log      core/Registration.cpp:93,137
//   SE3d * SE3d, inverse pipeline/sageICP.cpp:76,90,114,119
//   SE3d * Vector3d      core/Registration.cpp:106 (Eigen quaternion _transformVector form)
//   (the 6x6 ldlt().solve of core/Registration.cpp:92 is done through its 3x3 Schur complement in registration.cu)
#pragma once
#include <cmath>

#ifdef __CUDACC__
#define SAGE_HD __host__ __device__ __forceinline__
#else
#define SAGE_HD inline
#endif

namespace sage {

struct Pose {  // unit quaternion (w,x,y,z) + translation
    double qw, qx, qy, qz, tx, ty, tz;
};

SAGE_HD Pose pose_identity() { return {1, 0, 0, 0, 0, 0, 0}; }
SAGE_HD Pose pose_from_wire(const double *p) { return {p[6], p[3], p[4], p[5], p[0], p[1], p[2]}; }
SAGE_HD void pose_to_wire(const Pose &T, double *p) {
    p[0] = T.tx, p[1] = T.ty, p[2] = T.tz, p[3] = T.qx, p[4] = T.qy, p[5] = T.qz, p[6] = T.qw;
}

// p' = q * p + t with Eigen's _transformVector operation order: uv = 2 (qv x p); p + w uv + qv x uv.
// On device every operation is an explicit IEEE round-to-nearest op (no FMA contraction) so that the
// transformed coordinates, and therefore voxel keys and candidate ranking, are reproducible.
SAGE_HD void pose_act(const Pose &T, double x, double y, double z, double &ox, double &oy, double &oz) {
#ifdef __CUDA_ARCH__
    double ux = __dsub_rn(__dmul_rn(T.qy, z), __dmul_rn(T.qz, y));
    double uy = __dsub_rn(__dmul_rn(T.qz, x), __dmul_rn(T.qx, z));
    double uz = __dsub_rn(__dmul_rn(T.qx, y), __dmul_rn(T.qy, x));
    ux = __dadd_rn(ux, ux), uy = __dadd_rn(uy, uy), uz = __dadd_rn(uz, uz);
    const double cx = __dsub_rn(__dmul_rn(T.qy, uz), __dmul_rn(T.qz, uy));
    const double cy = __dsub_rn(__dmul_rn(T.qz, ux), __dmul_rn(T.qx, uz));
    const double cz = __dsub_rn(__dmul_rn(T.qx, uy), __dmul_rn(T.qy, ux));
    ox = __dadd_rn(__dadd_rn(__dadd_rn(x, __dmul_rn(T.qw, ux)), cx), T.tx);
    oy = __dadd_rn(__dadd_rn(__dadd_rn(y, __dmul_rn(T.qw, uy)), cy), T.ty);
    oz = __dadd_rn(__dadd_rn(__dadd_rn(z, __dmul_rn(T.qw, uz)), cz), T.tz);
#else
    double ux = T.qy * z - T.qz * y, uy = T.qz * x - T.qx * z, uz = T.qx * y - T.qy * x;
    ux += ux, uy += uy, uz += uz;
    const double cx = T.qy * uz - T.qz * uy, cy = T.qz * ux - T.qx * uz, cz = T.qx * uy - T.qy * ux;
    ox = ((x + T.qw * ux) + cx) + T.tx;
    oy = ((y + T.qw * uy) + cy) + T.ty;
    oz = ((z + T.qw * uz) + cz) + T.tz;
#endif
}

SAGE_HD Pose pose_mul(const Pose &a, const Pose &b) {
    Pose r;
    r.qw = a.qw * b.qw - a.qx * b.qx - a.qy * b.qy - a.qz * b.qz;
    r.qx = a.qw * b.qx + a.qx * b.qw + a.qy * b.qz - a.qz * b.qy;
    r.qy = a.qw * b.qy + a.qy * b.qw + a.qz * b.qx - a.qx * b.qz;
    r.qz = a.qw * b.qz + a.qz * b.qw + a.qx * b.qy - a.qy * b.qx;
    const double n = sqrt(r.qw * r.qw + r.qx * r.qx + r.qy * r.qy + r.qz * r.qz);
    r.qw /= n, r.qx /= n, r.qy /= n, r.qz /= n;
    Pose rot = a;
    rot.tx = rot.ty = rot.tz = 0;
    pose_act(rot, b.tx, b.ty, b.tz, r.tx, r.ty, r.tz);
    r.tx += a.tx, r.ty += a.ty, r.tz += a.tz;
    return r;
}

SAGE_HD Pose pose_inverse(const Pose &a) {
    Pose r{a.qw, -a.qx, -a.qy, -a.qz, 0, 0, 0};
    double x, y, z;
    pose_act(r, -a.tx, -a.ty, -a.tz, x, y, z);
    r.tx = x, r.ty = y, r.tz = z;
    return r;
}

// Sophus SE3::exp, tangent = (upsilon, omega)
SAGE_HD Pose pose_exp(const double xi[6]) {
    const double eps = 1e-10;
    const double wx = xi[3], wy = xi[4], wz = xi[5];
    const double th2 = (wx * wx + wy * wy) + wz * wz;
    double imag, real, theta;
    if (th2 < eps * eps) {
        theta = 0;
        const double th4 = th2 * th2;
        imag = 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * th4;
        real = 1.0 - (1.0 / 8.0) * th2 + (1.0 / 384.0) * th4;
    } else {
        theta = sqrt(th2);
        const double half = 0.5 * theta;
        imag = sin(half) / theta;
        real = cos(half);
    }
    Pose T;
    T.qw = real, T.qx = imag * wx, T.qy = imag * wy, T.qz = imag * wz;
    const double n = sqrt(T.qw * T.qw + T.qx * T.qx + T.qy * T.qy + T.qz * T.qz);
    T.qw /= n, T.qx /= n, T.qy /= n, T.qz /= n;
    // V = I + a*Om + b*Om^2, t = V * upsilon; Om*u = w x u, Om^2*u = w x (w x u)
    double a, b;
    if (theta < eps) {
        a = 0.5, b = 1.0 / 6.0;  // series limit of the closed form (Sophus uses R here; identical to O(theta))
    } else {
        a = (1.0 - cos(theta)) / th2;
        b = (theta - sin(theta)) / (th2 * theta);
    }
    const double ux = xi[0], uy = xi[1], uz = xi[2];
    const double c1x = wy * uz - wz * uy, c1y = wz * ux - wx * uz, c1z = wx * uy - wy * ux;
    const double c2x = wy * c1z - wz * c1y, c2y = wz * c1x - wx * c1z, c2z = wx * c1y - wy * c1x;
    T.tx = ux + a * c1x + b * c2x;
    T.ty = uy + a * c1y + b * c2y;
    T.tz = uz + a * c1z + b * c2z;
    return T;
}

// Sophus SE3::log -> (upsilon, omega)
SAGE_HD void pose_log(const Pose &T, double xi[6]) {
    const double eps = 1e-10;
    const double n2 = (T.qx * T.qx + T.qy * T.qy) + T.qz * T.qz;
    const double w = T.qw;
    double two_atan, theta;
    if (n2 < eps * eps) {
        two_atan = 2.0 / w - (2.0 / 3.0) * n2 / (w * w * w);
        theta = 2.0 * n2 / w;
    } else {
        const double n = sqrt(n2);
        const double at = (w < 0) ? atan2(-n, -w) : atan2(n, w);
        two_atan = 2.0 * at / n;
        theta = two_atan * n;
    }
    const double wx = two_atan * T.qx, wy = two_atan * T.qy, wz = two_atan * T.qz;
    double c;
    if (fabs(theta) < eps) {
        c = 1.0 / 12.0;
    } else {
        const double half = 0.5 * theta;
        c = (1.0 - theta * cos(half) / (2.0 * sin(half))) / (theta * theta);
    }
    // V^-1 t = t - 0.5 w x t + c w x (w x t)
    const double c1x = wy * T.tz - wz * T.ty, c1y = wz * T.tx - wx * T.tz, c1z = wx * T.ty - wy * T.tx;
    const double c2x = wy * c1z - wz * c1y, c2y = wz * c1x - wx * c1z, c2z = wx * c1y - wy * c1x;
    xi[0] = T.tx - 0.5 * c1x + c * c2x;
    xi[1] = T.ty - 0.5 * c1y + c * c2y;
    xi[2] = T.tz - 0.5 * c1z + c * c2z;
    xi[3] = wx, xi[4] = wy, xi[5] = wz;
}

// rotation angle in [0, pi]: Eigen::AngleAxisd(R).angle() (core/Threshold.cpp:30) = 2 atan2(|qv|, |qw|)
SAGE_HD double pose_rotation_angle(const Pose &T) {
    const double n = sqrt((T.qx * T.qx + T.qy * T.qy) + T.qz * T.qz);
    return (n != 0.0) ? 2.0 * atan2(n, fabs(T.qw)) : 0.0;
}

}  // namespace sage
