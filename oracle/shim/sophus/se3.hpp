// ORACLE — TEST INFRASTRUCTURE ONLY.  NOT Sophus.
//
// Stand-in for the part of Sophus::SE3d / SO3d (nachovizzo/Sophus 1.22.11, cpp/sage_icp/3rdparty/sophus/sophus.cmake:29; not in
// this image) that the reference's hot-path sources call, so that they compile unmodified into oracle/_ref (see Eigen/Core in
// this directory).  The arithmetic is the oracle's restatement of the published Sophus formulas (oracle/se3.hpp): what the
// reference build pins is therefore the reference's OWN code, not Sophus.
#pragma once
#include "../Eigen/Core"
#include "../Eigen/Geometry"

namespace Sophus {

using Vector6d = Eigen::Matrix<double, 6, 1>;

class SO3d {
public:
    static Eigen::Matrix3d hat(const Eigen::Vector3d &w) {
        Eigen::Matrix3d m = Eigen::Matrix3d::Zero();
        m(0, 1) = -w[2], m(0, 2) = w[1];
        m(1, 0) = w[2], m(1, 2) = -w[0];
        m(2, 0) = -w[1], m(2, 1) = w[0];
        return m;
    }
};

class SE3d {
public:
    SE3d() { T_.q = orc::Quat{1, 0, 0, 0}; }
    explicit SE3d(const orc::SE3 &T) : T_(T) {}
    const orc::SE3 &raw() const { return T_; }

    static SE3d exp(const Vector6d &xi) {
        const double a[6] = {xi[0], xi[1], xi[2], xi[3], xi[4], xi[5]};
        return SE3d(orc::se3_exp(a));
    }
    Vector6d log() const {
        double a[6];
        orc::se3_log(T_, a);
        Vector6d xi;
        for (int i = 0; i < 6; ++i) xi[i] = a[i];
        return xi;
    }
    SE3d inverse() const { return SE3d(orc::se3_inverse(T_)); }
    SE3d operator*(const SE3d &o) const { return SE3d(orc::se3_mul(T_, o.T_)); }
    Eigen::Vector3d operator*(const Eigen::Vector3d &p) const {
        const orc::Vec3 r = orc::se3_act(T_, orc::Vec3{p[0], p[1], p[2]});
        return Eigen::Vector3d(r.x, r.y, r.z);
    }
    Eigen::Vector3d translation() const { return Eigen::Vector3d(T_.t.x, T_.t.y, T_.t.z); }
    Eigen::Matrix3d rotationMatrix() const {
        const orc::Mat3 m = orc::quat_matrix(T_.q);
        Eigen::Matrix3d R;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) R(i, j) = m.m[i][j];
        return R;
    }

private:
    orc::SE3 T_;
};

}  // namespace Sophus
