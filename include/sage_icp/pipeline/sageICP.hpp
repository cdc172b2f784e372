// sage_icp::pipeline::sageICP — drop-in replacement header for the reference's cpp/sage_icp/pipeline/sageICP.hpp.
//
// Same include path ("sage_icp/pipeline/sageICP.hpp"), same namespace, same public surface (reference
// pipeline/sageICP.hpp:39-109): sageConfig with every field, both RegisterFrame overloads, Voxelize,
// GetAdaptiveThreshold, GetPredictionModel, HasMoved, TransformToLastFrame, LocalMap, poses, reinitialize — so
// ros/ros2/OdometryServer.{hpp,cpp} recompiles against it unchanged and links libsage_icp_b200.so instead of
// sage_icp::pipeline / sage_icp::core.  Everything below is a thin pimpl over the C ABI in sage_icp_b200.h; all compute
// runs in hand-written sm_100a CUDA on the handle's device.  There is no CPU fallback: constructing a sageICP from a
// config on a machine without a B200 throws std::runtime_error (the reference has no error path at all).
//
// Differences a maintainer should know:
//   * copies of a sageICP share one device pipeline (the reference deep-copies; the ROS node only move-assigns);
//   * LocalMap() returns the same point set in device block order, and the map drops every far voxel; after
//     SetReferenceMapSemantics(true) both follow the reference's tsl::robin_map (iteration order, erase-while-iterating sweep);
//   * the default-constructed object owns no device state until it is assigned from sageICP(config)
//     (the reference's default config has empty voxel_labels, which is undefined behaviour there: SURVEY.md A.11).
#pragma once

#include <Eigen/Core>
#include <Eigen/Geometry>
#include <memory>
#include <sophus/se3.hpp>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "sage_icp_b200.h"

namespace sage_icp::pipeline {

struct sageConfig {  // reference pipeline/sageICP.hpp:39-65, field for field
    std::vector<std::vector<int>> voxel_labels;
    std::vector<double> voxel_size;
    // map params
    double voxel_size_map = 1.0;

    double max_range = 100.0;
    double min_range = 5.0;
    double label_max_range = 50.0;
    double local_map_range = 100.0;
    int basic_points_per_voxel = 20;
    int critical_points_per_voxel = 20;
    std::vector<int> basic_parts_labels;

    // th parms
    double min_motion_th = 0.1;
    double initial_threshold = 2.0;
    double sem_th = 0.4;

    // Motion compensation
    bool deskew = false;

    bool dynamic_vehicle_filter = false;
    double dynamic_vehicle_filter_th = 0.5;
    int dynamic_vehicle_voxid = 5;
    std::vector<int> dynamic_remove_lankmark;

    // not in the reference: CUDA ordinal the pipeline lives on
    int device = 0;
};

class sageICP {
public:
    using Vector4dVector = std::vector<Eigen::Vector4d>;
    using Vector4dVectorTuple = std::tuple<Vector4dVector, double, double>;
    using Vector4dVectorTuple2 = std::tuple<Vector4dVector, Vector4dVector>;

    explicit sageICP(const sageConfig &config) : config_(config) {
        std::vector<int32_t> offsets{0}, labels;
        for (const auto &group : config.voxel_labels) {
            labels.insert(labels.end(), group.begin(), group.end());
            offsets.push_back(static_cast<int32_t>(labels.size()));
        }
        const std::vector<int32_t> basic(config.basic_parts_labels.begin(), config.basic_parts_labels.end());
        const std::vector<int32_t> landmark(config.dynamic_remove_lankmark.begin(), config.dynamic_remove_lankmark.end());
        sage_config_pod pod{};
        pod.n_groups = static_cast<int32_t>(config.voxel_size.size());
        pod.group_offsets = offsets.data();
        pod.group_labels = labels.data();
        pod.voxel_size = config.voxel_size.data();
        pod.voxel_size_map = config.voxel_size_map;
        pod.max_range = config.max_range;
        pod.min_range = config.min_range;
        pod.label_max_range = config.label_max_range;
        pod.local_map_range = config.local_map_range;
        pod.basic_points_per_voxel = config.basic_points_per_voxel;
        pod.critical_points_per_voxel = config.critical_points_per_voxel;
        pod.n_basic_parts_labels = static_cast<int32_t>(basic.size());
        pod.basic_parts_labels = basic.data();
        pod.min_motion_th = config.min_motion_th;
        pod.initial_threshold = config.initial_threshold;
        pod.sem_th = config.sem_th;
        pod.deskew = config.deskew ? 1 : 0;
        pod.dynamic_vehicle_filter = config.dynamic_vehicle_filter ? 1 : 0;
        pod.dynamic_vehicle_filter_th = config.dynamic_vehicle_filter_th;
        pod.dynamic_vehicle_voxid = config.dynamic_vehicle_voxid;
        pod.n_dynamic_remove_lankmark = static_cast<int32_t>(landmark.size());
        pod.dynamic_remove_lankmark = landmark.data();
        if (config.voxel_labels.size() != config.voxel_size.size())
            throw std::invalid_argument("sageConfig: voxel_labels and voxel_size differ in length");
        sage_pipeline *h = sage_create(&pod, config.device);
        if (!h) throw std::runtime_error(std::string("sage_icp_b200: ") + sage_last_error());
        handle_.reset(h, [](sage_pipeline *p) { sage_destroy(p); });
    }

    sageICP() = default;  // no device state; assign from sageICP(config) before use (ros/ros2/OdometryServer.cpp:104)

    // pipeline/sageICP.cpp:54-95
    Vector4dVectorTuple RegisterFrame(const std::vector<Eigen::Vector4d> &frame) { return Register(frame, nullptr); }
    // pipeline/sageICP.cpp:36-52
    Vector4dVectorTuple RegisterFrame(const std::vector<Eigen::Vector4d> &frame, const std::vector<double> &timestamps) {
        return Register(frame, timestamps.size() == frame.size() && !timestamps.empty() ? timestamps.data() : nullptr);
    }
    // Extension (not in the reference): RegisterFrame straight from a sensor_msgs/PointCloud2 data buffer, i.e.
    // utils::PointCloud2ToEigen (ros/ros2/Utils.hpp:161-180) + RegisterFrame with the f32 -> f64 widening done on the
    // device.  In OdometryServer::RegisterFrame (ros/ros2/OdometryServer.cpp:156-167) this replaces the two calls
    //   const auto points = utils::PointCloud2ToEigen(msg);  odometry_.RegisterFrame(points, timestamps);
    // by  odometry_.RegisterFramePointCloud2(msg->data.data(), msg->width * msg->height, msg->point_step, 0, 4, 8, 12,
    //                                        msg->fields.size() == 5 ? 2 : 7, timestamps);
    Vector4dVectorTuple RegisterFramePointCloud2(const uint8_t *data, size_t n_points, uint32_t point_step, uint32_t x_offset,
                                                 uint32_t y_offset, uint32_t z_offset, uint32_t label_offset, int label_datatype,
                                                 const std::vector<double> &timestamps) {
        double pose[7], t_icp = 0, t_all = 0;
        Check(sage_register_frame_pointcloud2(Handle(), data, n_points, point_step, x_offset, y_offset, z_offset, label_offset,
                                              label_datatype, timestamps.size() == n_points && n_points ? timestamps.data() : nullptr,
                                              pose, &t_icp, &t_all));
        return {FetchSource(), t_icp, t_all};
    }
    // pipeline/sageICP.cpp:97-101 — returns {source, frame_downsample}
    Vector4dVectorTuple2 Voxelize(const std::vector<Eigen::Vector4d> &frame) const {
        Vector4dVector source(frame.size()), downsample(frame.size());
        size_t n_source = 0, n_downsample = 0;
        Check(sage_voxelize(Handle(), Data(frame), frame.size(), Data(source), &n_source, Data(downsample), &n_downsample));
        source.resize(n_source);
        downsample.resize(n_downsample);
        return {std::move(source), std::move(downsample)};
    }
    double GetAdaptiveThreshold() { return sage_get_adaptive_threshold(Handle()); }  // pipeline/sageICP.cpp:103-108
    Sophus::SE3d GetPredictionModel() const {                                         // pipeline/sageICP.cpp:110-115
        double p[7];
        Check(sage_get_prediction_model(Handle(), p));
        return ToSE3(p);
    }
    bool HasMoved() { return sage_has_moved(Handle()) != 0; }  // pipeline/sageICP.cpp:117-121
    // pipeline/sageICP.cpp:123-129
    std::vector<Eigen::Vector4d> TransformToLastFrame(const Sophus::SE3d &last_pose, const Sophus::SE3d &current_pose,
                                                      const std::vector<Eigen::Vector4d> &points) {
        double a[7], b[7];
        FromSE3(last_pose, a);
        FromSE3(current_pose, b);
        Vector4dVector out(points.size());
        Check(sage_transform_to_last_frame(handle_.get(), a, b, Data(points), points.size(), Data(out)));
        return out;
    }

    // Extra C++ API to facilitate ROS debugging (pipeline/sageICP.hpp:91-99)
    std::vector<Eigen::Vector4d> LocalMap() const {
        const int64_t n = sage_local_map(Handle(), nullptr, 0);
        if (n < 0) Check(static_cast<int>(n));
        Vector4dVector out(static_cast<size_t>(n));
        const int64_t k = sage_local_map(Handle(), Data(out), out.size());
        if (k < 0) Check(static_cast<int>(k));
        out.resize(static_cast<size_t>(k));
        return out;
    }
    // The node calls poses() after every scan (ros/ros2/OdometryServer.cpp:173).  Poses never change once pushed, so the adaptor
    // keeps the ones it has converted and fetches only the new tail, in ONE bulk C-ABI call (sage_get_poses): O(1) per frame
    // instead of O(frames) calls.  reinitialize() (fewer poses than cached) drops the cache.
    std::vector<Sophus::SE3d> poses() const {
        if (!handle_) return {};
        const int64_t n = sage_num_poses(handle_.get());
        if (n < 0) Check(static_cast<int>(n));
        if (static_cast<size_t>(n) < pose_cache_.size()) pose_cache_.clear();
        const size_t have = pose_cache_.size(), want = static_cast<size_t>(n) - have;
        if (want > 0) {
            std::vector<double> raw(7 * want);
            const int64_t got = sage_get_poses(handle_.get(), have, raw.data(), want);
            if (got < 0) Check(static_cast<int>(got));
            pose_cache_.reserve(static_cast<size_t>(n));
            for (int64_t i = 0; i < got; ++i) pose_cache_.push_back(ToSE3(raw.data() + 7 * i));
        }
        return pose_cache_;
    }
    bool reinitialize() {
        if (handle_) Check(sage_reset(handle_.get()));
        pose_cache_.clear();
        return true;
    }
    // Extension (not in the reference): true = reproduce the reference map's tsl::robin_map behaviour — the far voxels its
    // erase-while-iterating sweep skips (core/VoxelHashMap.cpp:176-184) and LocalMap() in its iteration order.  Call right after
    // construction or reinitialize() (the map must be empty).  Costs a few host round trips per frame.
    // Extension (not in the reference, which is CPU only): run this object on the given CUDA devices.  One id: that GPU.  Several:
    // the ICP queries of every RegisterFrame are sharded over them (replicated map, sums all-reduced inside the search kernel over
    // NVLink peer memory) — the one line a node adds after constructing sageICP to use a multi-GPU box.  Fresh object only.
    void SetDevices(const std::vector<int> &cuda_devices) {
        Check(sage_set_devices(Handle(), cuda_devices.data(), static_cast<int>(cuda_devices.size())));
    }
    void SetReferenceMapSemantics(bool on) { Check(sage_map_set_eviction(sage_pipeline_map(Handle()), on ? 1 : 0)); }

private:
    static double *Data(Vector4dVector &v) { return v.empty() ? nullptr : v.front().data(); }
    static const double *Data(const Vector4dVector &v) { return v.empty() ? nullptr : v.front().data(); }
    static Sophus::SE3d ToSE3(const double p[7]) {
        return Sophus::SE3d(Eigen::Quaterniond(p[6], p[3], p[4], p[5]), Eigen::Vector3d(p[0], p[1], p[2]));
    }
    static void FromSE3(const Sophus::SE3d &T, double p[7]) {
        const auto t = T.translation();
        const auto q = T.unit_quaternion();
        p[0] = t[0], p[1] = t[1], p[2] = t[2], p[3] = q.x(), p[4] = q.y(), p[5] = q.z(), p[6] = q.w();
    }
    static void Check(int rc) {
        if (rc < 0) throw std::runtime_error(std::string("sage_icp_b200: ") + sage_last_error());
    }
    sage_pipeline *Handle() const {
        if (!handle_) throw std::logic_error("sageICP used before it was constructed from a sageConfig");
        return handle_.get();
    }
    Vector4dVectorTuple Register(const Vector4dVector &frame, const double *timestamps) {
        static_assert(sizeof(Eigen::Vector4d) == 4 * sizeof(double), "Vector4d must be 4 packed doubles");
        double pose[7], t_icp = 0, t_all = 0;
        Check(sage_register_frame(Handle(), Data(frame), frame.size(), timestamps, pose, &t_icp, &t_all));
        return {FetchSource(), t_icp, t_all};
    }
    Vector4dVector FetchSource() {
        const int64_t n = sage_last_source(Handle(), nullptr, 0);
        Vector4dVector source(static_cast<size_t>(n > 0 ? n : 0));
        if (n > 0) sage_last_source(Handle(), Data(source), source.size());
        return source;
    }

    sageConfig config_;
    std::shared_ptr<sage_pipeline> handle_;
    mutable std::vector<Sophus::SE3d> pose_cache_;  // poses() already converted (copies of a sageICP share the handle, each keeps its own)
};

}  // namespace sage_icp::pipeline
