// Per-frame orchestration: sage_icp::pipeline::sageICP (pipeline/sageICP.hpp:67-109, pipeline/sageICP.cpp:36-129)
// and sage_icp::AdaptiveThreshold (core/Threshold.hpp:29-52, core/Threshold.cpp:29-50).  Host C++ drives the
// device front end, the device-resident Gauss-Newton loop and the device map update; the scan crosses PCIe once
// in and the pose once out.
#include "pipeline.cuh"

#include <chrono>

namespace sage {

static GroupTable make_groups(const sage_config_pod &c) {
    if (c.n_groups < 1) throw ArgError("sageConfig needs at least one voxel group (voxel_labels / voxel_size)");
    if (c.n_groups > kMaxGroups) throw ArgError("at most 16 voxel groups supported");
    if (!c.voxel_size || !c.group_offsets || !c.group_labels) throw ArgError("sageConfig: voxel_size / group_offsets / group_labels is NULL");
    if (c.group_offsets[0] != 0) throw ArgError("group_offsets must start at 0");
    GroupTable g{};
    g.n_groups = c.n_groups;
    g.n_labels = 0;
    for (int i = 0; i < c.n_groups; ++i) {
        g.voxel_size[i] = c.voxel_size[i];
        if (!(c.voxel_size[i] > 0)) throw ArgError("voxel_size must be positive");
        if (c.group_offsets[i + 1] < c.group_offsets[i]) throw ArgError("group_offsets must be non-decreasing");
        for (int k = c.group_offsets[i]; k < c.group_offsets[i + 1]; ++k) {
            if (g.n_labels >= kMaxGroupLabels) throw ArgError("at most 64 labels across voxel groups supported");
            g.label[g.n_labels] = c.group_labels[k];
            g.group_of[g.n_labels] = i;
            ++g.n_labels;
        }
    }
    return g;
}

Pipeline::Pipeline(const sage_config_pod &c, int device)
    : cfg_(c),
      map_(c.voxel_size_map, c.local_map_range, c.basic_points_per_voxel, c.critical_points_per_voxel, c.basic_parts_labels,
           c.n_basic_parts_labels, device),
      fe_(make_groups(c), device, map_.stream()) {
    // the reference evaluates voxel_labels[dynamic_vehicle_voxid] unconditionally (pipeline/sageICP.cpp:63): out of range
    // there is UB, here it is an error
    if (c.dynamic_vehicle_voxid < 0 || c.dynamic_vehicle_voxid >= c.n_groups) throw ArgError("dynamic_vehicle_voxid out of range");
    {  // dynamic-vehicle branch of Preprocess: labels of the vehicle voxel group and the landmark labels (pipeline/sageICP.cpp:61-64)
        const int g = c.dynamic_vehicle_voxid;
        const int n_dyn = c.group_offsets[g + 1] - c.group_offsets[g];
        if (n_dyn > 32 || c.n_dynamic_remove_lankmark > 32) throw ArgError("at most 32 dynamic / landmark labels supported");
        dyn_.n_dynamic = n_dyn;
        for (int k = 0; k < n_dyn; ++k) dyn_.dynamic_labels[k] = c.group_labels[c.group_offsets[g] + k];
        dyn_.n_landmark = c.n_dynamic_remove_lankmark;
        for (int k = 0; k < c.n_dynamic_remove_lankmark; ++k) dyn_.landmark_labels[k] = c.dynamic_remove_lankmark[k];
        dyn_.dy_th = c.dynamic_vehicle_filter_th;
    }
    // own copies of the label arrays (the POD's pointers belong to the caller)
    cfg_.group_offsets = cfg_.group_labels = cfg_.basic_parts_labels = cfg_.dynamic_remove_lankmark = nullptr;
    cfg_.voxel_size = nullptr;
    reset_threshold();
}

void Pipeline::reset_threshold() {
    model_error_sse2_ = 0;
    num_samples_ = 0;
    model_deviation_ = pose_identity();
}

void Pipeline::key_frame_grid(const double *xyzl, size_t n, const Pose *last, const Pose *current, const double bounds[6], int rows, int cols,
                              const int32_t *last_occ, int32_t *grid_out, double *overlap) {
    if (rows < 1 || cols < 1 || (long long)rows * cols > (1ll << 26)) throw ArgError("key-frame grid: rows/cols out of range");
    OccGridParams g;
    g.x0 = bounds[0], g.x1 = bounds[1], g.y0 = bounds[2], g.y1 = bounds[3], g.z0 = bounds[4], g.z1 = bounds[5];
    g.rows = rows, g.cols = cols;
    g.x_res = (g.x1 - g.x0) / cols;  // Width  (Utils.hpp:224)
    g.y_res = (g.y1 - g.y0) / rows;  // Height (Utils.hpp:225)
    Pose T;
    const bool move = last && current;
    if (move) T = pose_mul(pose_inverse(*last), *current);  // last_pose.inverse() * current_pose, pipeline/sageICP.cpp:127
    fe_.key_frame_grid(map_.stage_points(xyzl, n), n, move ? &T : nullptr, g, last_occ, grid_out, overlap);
}

void Pipeline::reinitialize() {  // pipeline/sageICP.hpp:94-99
    poses_.clear();
    reset_threshold();
    map_.clear();
}

// AdaptiveThreshold::ComputeThreshold — core/Threshold.cpp:39-50
double Pipeline::compute_threshold() {
    const double theta = pose_rotation_angle(model_deviation_);
    const double delta_rot = 2.0 * cfg_.max_range * std::sin(theta / 2.0);
    const double delta_trans = std::sqrt((model_deviation_.tx * model_deviation_.tx + model_deviation_.ty * model_deviation_.ty) +
                                         model_deviation_.tz * model_deviation_.tz);
    const double model_error = delta_trans + delta_rot;
    if (model_error > cfg_.min_motion_th) {
        model_error_sse2_ += model_error * model_error;
        num_samples_++;
    }
    if (num_samples_ < 1) return cfg_.initial_threshold;
    return std::sqrt(model_error_sse2_ / num_samples_);
}

bool Pipeline::has_moved() {  // pipeline/sageICP.cpp:117-121
    if (poses_.empty()) return false;
    const Pose d = pose_mul(pose_inverse(poses_.front()), poses_.back());
    const double motion = std::sqrt((d.tx * d.tx + d.ty * d.ty) + d.tz * d.tz);
    return motion > 5.0 * cfg_.min_motion_th;
}

double Pipeline::get_adaptive_threshold() {  // pipeline/sageICP.cpp:103-108
    if (!has_moved()) return cfg_.initial_threshold;
    return compute_threshold();
}

Pose Pipeline::get_prediction_model() const {  // pipeline/sageICP.cpp:110-115
    const size_t N = poses_.size();
    if (N < 2) return pose_identity();
    return pose_mul(pose_inverse(poses_[N - 2]), poses_[N - 1]);
}

CropParams Pipeline::crop() const { return CropParams{1, cfg_.max_range, cfg_.min_range, cfg_.label_max_range}; }

// Voxelize — pipeline/sageICP.cpp:97-101 (device in, device out)
void Pipeline::voxelize_dev(const double4 *frame, size_t n, const CropParams &cp) {
    ds_.ensure(n ? n : 1);
    src_.ensure(n ? n : 1);
    n_ds_ = fe_.downsample(frame, n, 0.5, cp, ds_.p);
    n_src_ = fe_.downsample(ds_.p, n_ds_, 1.5, CropParams{0, 0, 0, 0}, src_.p);
}

void Pipeline::voxelize_host(const double *xyzl, size_t n, std::vector<double> &source, std::vector<double> &downsample) {
    double4 *raw = map_.stage_points(xyzl, n);
    voxelize_dev(raw, n, CropParams{0, 0, 0, 0});
    fetch(src_.p, n_src_, source);
    fetch(ds_.p, n_ds_, downsample);
}

void Pipeline::fetch(const double4 *dev, size_t n, std::vector<double> &out) {
    out.resize(n * 4);
    if (n) {
        SAGE_CUDA(cudaMemcpyAsync(out.data(), dev, n * sizeof(double4), cudaMemcpyDeviceToHost, map_.stream()));
        SAGE_CUDA(cudaStreamSynchronize(map_.stream()));
    }
}

// RegisterFrame — pipeline/sageICP.cpp:36-52 (deskew wrapper) and :54-95
void Pipeline::register_frame(const double *xyzl, size_t n, const double *timestamps, Pose &pose_out, double &t_icp, double &t_all) {
    register_frame_dev(map_.stage_points(xyzl, n), n, timestamps, pose_out, t_icp, t_all);
}

void Pipeline::register_frame_pointcloud2(const uint8_t *data, size_t n, uint32_t point_step, uint32_t x_off, uint32_t y_off, uint32_t z_off,
                                          uint32_t label_off, int label_is_f32, const double *timestamps, Pose &pose_out, double &t_icp,
                                          double &t_all) {
    const uint32_t need = (label_is_f32 ? 4u : 1u);
    if (point_step == 0 || x_off + 4 > point_step || y_off + 4 > point_step || z_off + 4 > point_step || label_off + need > point_step)
        throw ArgError("PointCloud2 field offsets do not fit point_step");
    SAGE_CUDA(cudaSetDevice(map_.device()));
    packed_.ensure(n * (size_t)point_step + 1);
    unpacked_.ensure(n ? n : 1);
    if (n) SAGE_CUDA(cudaMemcpyAsync(packed_.p, data, n * (size_t)point_step, cudaMemcpyHostToDevice, map_.stream()));
    fe_.unpack_pointcloud2(packed_.p, n, point_step, x_off, y_off, z_off, label_off, label_is_f32, unpacked_.p);
    register_frame_dev(unpacked_.p, n, timestamps, pose_out, t_icp, t_all);
}

void Pipeline::register_frame_dev(const double4 *raw, size_t n, const double *timestamps, Pose &pose_out, double &t_icp, double &t_all) {
    using clock = std::chrono::high_resolution_clock;
    if (cfg_.deskew && timestamps && poses_.size() > 2) {  // pipeline/sageICP.cpp:39-49
        ts_.ensure(n ? n : 1);
        deskewed_.ensure(n ? n : 1);
        if (n) SAGE_CUDA(cudaMemcpyAsync(ts_.p, timestamps, n * sizeof(double), cudaMemcpyHostToDevice, map_.stream()));
        const size_t N = poses_.size();
        fe_.deskew(raw, ts_.p, n, poses_[N - 2], poses_[N - 1], deskewed_.p);
        raw = deskewed_.p;
    }
    const auto t0 = clock::now();
    if (cfg_.dynamic_vehicle_filter) {  // Preprocess with the vehicle filter, then Voxelize
        filtered_.ensure(n ? n : 1);
        const size_t nf = preprocess_dev(raw, n, filtered_.p);
        voxelize_dev(filtered_.p, nf, CropParams{0, 0, 0, 0});
    } else {
        voxelize_dev(raw, n, crop());  // Preprocess (range branch) fused into the first downsample pass
    }
    const double sigma = get_adaptive_threshold();
    const Pose prediction = get_prediction_model();
    const Pose last_pose = !poses_.empty() ? poses_.back() : pose_identity();
    const Pose initial_guess = pose_mul(last_pose, prediction);
    const auto t1 = clock::now();
    Pose new_pose;
    last_iters_ = map_.register_frame_dev(src_.p, n_src_, initial_guess, 3.0 * sigma, sigma / 3.0, cfg_.sem_th, 500, 1e-4, new_pose);
    const auto t2 = clock::now();
    model_deviation_ = pose_mul(pose_inverse(initial_guess), new_pose);
    map_.update_dev(ds_.p, n_ds_, new_pose);  // asynchronous: overlaps the caller and the next frame's upload
    poses_.push_back(new_pose);
    last_sigma_ = sigma;
    pose_out = new_pose;
    t_icp = std::chrono::duration<double>(t2 - t1).count();
    t_all = std::chrono::duration<double>(t2 - t0).count();
}

size_t Pipeline::preprocess_dev(const double4 *raw, size_t n, double4 *out) {
    return cfg_.dynamic_vehicle_filter ? fe_.preprocess_dynamic(raw, n, crop(), dyn_, out) : fe_.preprocess(raw, n, crop(), out);
}

long long Pipeline::preprocess_host(const double *xyzl, size_t n, std::vector<double> &out) {
    double4 *raw = map_.stage_points(xyzl, n);
    tmp_.ensure(n ? n : 1);
    const size_t m = preprocess_dev(raw, n, tmp_.p);
    fetch(tmp_.p, m, out);
    return (long long)m;
}

long long Pipeline::downsample_host(const double *xyzl, size_t n, double scale, std::vector<double> &out) {
    double4 *raw = map_.stage_points(xyzl, n);
    tmp_.ensure(n ? n : 1);
    const size_t m = fe_.downsample(raw, n, scale, CropParams{0, 0, 0, 0}, tmp_.p);
    fetch(tmp_.p, m, out);
    return (long long)m;
}

}  // namespace sage
