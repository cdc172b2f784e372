// TEST DOUBLE — see Eigen/Core in this directory.  Only construction from (quaternion, translation) and the two accessors.
#pragma once
#include <Eigen/Core>
#include <Eigen/Geometry>
namespace Sophus {
struct SE3d {
    Eigen::Quaterniond q{1, 0, 0, 0};
    Eigen::Vector3d t;
    SE3d() = default;
    SE3d(const Eigen::Quaterniond &q_, const Eigen::Vector3d &t_) : q(q_), t(t_) {}
    const Eigen::Quaterniond &unit_quaternion() const { return q; }
    const Eigen::Vector3d &translation() const { return t; }
};
}  // namespace Sophus
