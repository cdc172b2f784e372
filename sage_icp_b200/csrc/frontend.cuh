// Scan front end on the device: range crop + per-label-group voxel downsample with the reference's output order.
// Mirrors sage_icp::Preprocess (range branch, core/Preprocessing.cpp:173-187), sage_icp::VoxelDownsample
// (core/Preprocessing.cpp:44-84) and sage_icp::DeSkewScan (core/Deskew.cpp:36-50).
#pragma once
#include <vector>

#include "common.cuh"

namespace sage {

constexpr int kMaxGroups = 16;
constexpr int kMaxGroupLabels = 64;

struct GroupTable {  // passed to kernels by value
    int n_groups;
    int n_labels;
    int label[kMaxGroupLabels];
    int group_of[kMaxGroupLabels];  // parallel to label[], in first-match order
    double voxel_size[kMaxGroups];
};

struct DynFilterParams {  // Preprocess' dynamic-vehicle branch (core/Preprocessing.cpp:95-172)
    int n_dynamic;
    int dynamic_labels[32];  // voxel_labels[dynamic_vehicle_voxid]
    int n_landmark;
    int landmark_labels[32];  // dynamic_remove_lankmark
    double dy_th;             // dynamic_vehicle_filter_th
};

struct OccGridParams {  // utils::EigenToGridMap (ros/ros2/Utils.hpp:220-242): bounds [[x0,x1],[y0,y1],[z0,z1]], H x W cells
    double x0, x1, y0, y1, z0, z1;
    double x_res, y_res;  // (x1 - x0) / cols, (y1 - y0) / rows
    int rows, cols;
};

struct CropParams {
    int enabled;
    double max_range, min_range, label_max_range;
};

// Iteration order of a tsl::robin_map v1.0.1 that received `n` DISTINCT keys with the given 20-bit hashes in this
// order.  order_out[j] = input position of the j-th element in iteration order.
void robin_iteration_order(const uint32_t *hash20, size_t n, uint32_t *order_out);

class FrontEnd {
public:
    FrontEnd(const GroupTable &groups, int device, cudaStream_t stream);

    // VoxelDownsample (optionally fused with the range crop) of a device-resident cloud.  Writes the result in the
    // reference's order to `out` (device, capacity >= n) and returns the number of points kept.  Synchronises.
    size_t downsample(const double4 *in, size_t n, double vox_scale, const CropParams &crop, double4 *out);
    // Preprocess only (order-preserving compaction).  Returns kept count.  Synchronises.
    size_t preprocess(const double4 *in, size_t n, const CropParams &crop, double4 *out);
    // Preprocess with dynamic_vehicle_filter = true: range crop, then vehicle-labelled points survive only in clusters
    // (0.5 m single linkage, >= 5 points) that touch enough landmark-labelled points.  Non-vehicle inliers first (input
    // order), then the kept vehicle points cluster by cluster in the reference's order (clusters by descending size, then by
    // smallest member; members ascending — core/Preprocessing.cpp:141-170 over PCL's cluster order).  Returns the kept
    // count.  Synchronises.
    size_t preprocess_dynamic(const double4 *in, size_t n, const CropParams &crop, const DynFilterParams &dyn, double4 *out);
    // utils::PointCloud2ToEigen (ros/ros2/Utils.hpp:161-180) on the device: packed records -> x, y, z, label as f64
    void unpack_pointcloud2(const uint8_t *data_dev, size_t n, uint32_t point_step, uint32_t x_off, uint32_t y_off, uint32_t z_off,
                            uint32_t label_off, int label_is_f32, double4 *out);
    // Key-frame occupancy grid (+ overlap against `last_occ_host` when given) of a device-resident cloud; T: transform applied first.
    void key_frame_grid(const double4 *pts, size_t n, const Pose *T, const OccGridParams &g, const int32_t *last_occ_host, int32_t *grid_host,
                        double *overlap);
    // DeSkewScan on the device (in place allowed).
    void deskew(const double4 *in, const double *timestamps_dev, size_t n, const Pose &start, const Pose &finish, double4 *out);

private:
    void scan_flags(size_t n, uint32_t *total_out, const uint32_t *err);  // exclusive scan of flags_ -> pos_, total -> *total_out (device or mapped pinned)

    GroupTable groups_;
    int device_;
    cudaStream_t stream_;
    DevBuf<unsigned long long> tkey_;
    DevBuf<uint32_t> tfirst_;
    uint32_t tcap_ = 0;
    DevBuf<uint32_t> slot_, flags_, pos_, block_sums_, widx_;
    DevBuf<uint32_t> total_;
    // dynamic-vehicle filter scratch
    DevBuf<unsigned long long> cell_key_;
    DevBuf<uint32_t> cell_head_v_, cell_head_l_, next_v_, next_l_, parent_, csize_, clm_, cls_;
    uint32_t cell_cap_ = 0;
    DevBuf<unsigned long long> dyn_key_[2];  // cluster-order sort of the kept vehicle points
    DevBuf<uint32_t> dyn_val_[2];
    DevBuf<uint8_t> sort_tmp_;
    DevBuf<int32_t> occ_;  // key-frame grids: current, last, two counters
    PinBuf<uint32_t> total_pin_, whash_pin_, perm_pin_[2];
    int parity_ = 0;
    // SAGE_FE_TRACE=1: host-side time of VoxelDownsample by part, printed when the front end is destroyed (tools/stream_bench.py)
    bool trace_ = false;
    double t_enqueue_ = 0, t_wait_ = 0, t_group_ = 0, t_replay_ = 0, t_tail_ = 0;
    long long n_calls_ = 0, n_keys_ = 0;
public:
    ~FrontEnd();
private:
    std::vector<std::vector<uint32_t>> group_members_, group_hashes_, group_order_;
};

}  // namespace sage
