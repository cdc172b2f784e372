// Host-side mirror of sage_icp::pipeline::sageICP (pipeline/sageICP.hpp:67-109) over the device kernels.
#pragma once
#include <memory>
#include <vector>

#include "../../include/sage_icp_b200.h"
#include "frontend.cuh"
#include "voxel_map.cuh"

namespace sage {

class Pipeline {
public:
    Pipeline(const sage_config_pod &config, int device);

    void register_frame(const double *xyzl, size_t n, const double *timestamps, Pose &pose_out, double &t_icp, double &t_all);
    // RegisterFrame fed with the raw sensor_msgs/PointCloud2 data buffer (host): unpacked on the device
    void register_frame_pointcloud2(const uint8_t *data, size_t n, uint32_t point_step, uint32_t x_off, uint32_t y_off, uint32_t z_off,
                                    uint32_t label_off, int label_is_f32, const double *timestamps, Pose &pose_out, double &t_icp,
                                    double &t_all);
    void voxelize_host(const double *xyzl, size_t n, std::vector<double> &source, std::vector<double> &downsample);
    double get_adaptive_threshold();
    bool has_moved();
    Pose get_prediction_model() const;
    void reinitialize();
    // sage_set_devices with n > 1: this process drives one map replica per extra GPU.  The front end runs on the first device;
    // the ICP queries are cut into contiguous shards, one per GPU, registered concurrently (one host thread per replica) with the
    // 17 sums all-reduced inside the search kernel over peer memory; every replica applies the same map update.  Only on a fresh
    // or reinitialised pipeline.
    void set_replica_devices(const std::vector<int> &devices);
    size_t n_devices() const { return replicas_.size() + 1; }
    long long preprocess_host(const double *xyzl, size_t n, std::vector<double> &out);
    long long downsample_host(const double *xyzl, size_t n, double scale, std::vector<double> &out);

    // the node's key-frame test (ros/ros2/OdometryServer.cpp:222-241): grid of `points` (moved by last^-1 * current when both are
    // given) and, when last_occ is given, its overlap with that grid
    void key_frame_grid(const double *xyzl, size_t n, const Pose *last, const Pose *current, const double bounds[6], int rows, int cols,
                        const int32_t *last_occ, int32_t *grid_out, double *overlap);

    const std::vector<Pose> &poses() const { return poses_; }
    VoxelMapGPU &map() { return map_; }
    void last_source(std::vector<double> &out) { fetch(src_.p, n_src_, out); }
    void last_downsample(std::vector<double> &out) { fetch(ds_.p, n_ds_, out); }
    size_t n_source() const { return n_src_; }
    size_t n_downsample() const { return n_ds_; }
    int last_iterations() const { return last_iters_; }
    double last_sigma() const { return last_sigma_; }

private:
    void voxelize_dev(const double4 *frame, size_t n, const CropParams &cp);
    void register_frame_dev(const double4 *raw, size_t n, const double *timestamps, Pose &pose_out, double &t_icp, double &t_all);
    void fetch(const double4 *dev, size_t n, std::vector<double> &out);
    double compute_threshold();
    void reset_threshold();
    CropParams crop() const;
    size_t preprocess_dev(const double4 *raw, size_t n, double4 *out);  // Preprocess, either branch
    DynFilterParams dyn_{};

    struct Replica {
        int device;
        std::unique_ptr<VoxelMapGPU> map;
        DevBuf<double4> src, ds;
    };
    std::vector<std::unique_ptr<Replica>> replicas_;
    std::vector<int> basic_labels_;  // kept for building replicas
    int register_sharded(const Pose &guess, double max_dist, double kernel, Pose &pose_out);  // ICP + map update on every GPU

    sage_config_pod cfg_;  // scalar fields only (array pointers nulled)
    VoxelMapGPU map_;
    FrontEnd fe_;
    std::vector<Pose> poses_;
    // AdaptiveThreshold state (core/Threshold.hpp:46-51)
    double model_error_sse2_ = 0;
    int num_samples_ = 0;
    Pose model_deviation_ = pose_identity();

    DevBuf<double4> ds_, src_, tmp_, deskewed_, filtered_;
    DevBuf<double> ts_;
    DevBuf<uint8_t> packed_;
    DevBuf<double4> unpacked_;
    size_t n_ds_ = 0, n_src_ = 0;
    int last_iters_ = 0;
    double last_sigma_ = 0;
};

}  // namespace sage
