// Per-frame orchestration: sage_icp::pipeline::sageICP (pipeline/sageICP.hpp:67-109, pipeline/sageICP.cpp:36-129)
// and sage_icp::AdaptiveThreshold (core/Threshold.hpp:29-52, core/Threshold.cpp:29-50).  Host C++ drives the
// device front end, the device-resident Gauss-Newton loop and the device map update; the scan crosses PCIe once
// in and the pose once out.
#include "pipeline.cuh"

#include <chrono>
#include <cstring>
#include <exception>
#include <thread>

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
    basic_labels_.assign(c.basic_parts_labels, c.basic_parts_labels + c.n_basic_parts_labels);
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
    for (auto &r : replicas_) r->map->clear();
}

void Pipeline::set_replica_devices(const std::vector<int> &devices) {
    if (!poses_.empty() || !map_.empty()) throw ArgError("the devices can only be changed on a fresh or reinitialised pipeline");
    if (devices.size() + 1 > 8) throw ArgError("at most 8 GPUs of one node");
    for (size_t i = 0; i < devices.size(); ++i) {
        if (devices[i] == map_.device()) throw ArgError("sage_set_devices: a GPU is listed twice");
        for (size_t j = 0; j < i; ++j)
            if (devices[i] == devices[j]) throw ArgError("sage_set_devices: a GPU is listed twice");
    }
    map_.peer_detach();
    replicas_.clear();
    if (devices.empty()) return;
    for (int d : devices) {
        std::unique_ptr<Replica> r(new Replica);
        r->device = d;
        r->map.reset(new VoxelMapGPU(cfg_.voxel_size_map, cfg_.local_map_range, cfg_.basic_points_per_voxel, cfg_.critical_points_per_voxel,
                                     basic_labels_.data(), (int)basic_labels_.size(), d));
        r->map->set_eviction_faithful(map_.eviction_faithful());
        replicas_.push_back(std::move(r));
    }
    const int world = (int)replicas_.size() + 1;
    double *bufs[8];
    int devs[8];
    bufs[0] = map_.peer_local_buffer(), devs[0] = map_.device();
    for (int k = 1; k < world; ++k) bufs[k] = replicas_[k - 1]->map->peer_local_buffer(), devs[k] = replicas_[k - 1]->device;
    map_.peer_attach_local(0, world, bufs, devs);
    for (int k = 1; k < world; ++k) replicas_[k - 1]->map->peer_attach_local(k, world, bufs, devs);
}

// sage_icp::RegisterFrame + VoxelHashMap::Update on every GPU of the handle.  Rank r registers the contiguous query slice
// [n r / R, n (r + 1) / R) against its own replica of the map; the sums meet inside the search kernel, so every rank takes the same
// steps and returns the same pose, bit for bit; then every rank applies the same update to its replica.
int Pipeline::register_sharded(const Pose &guess, double max_dist, double kernel, Pose &pose_out) {
    const int world = (int)replicas_.size() + 1;
    SAGE_CUDA(cudaSetDevice(map_.device()));
    SAGE_CUDA(cudaStreamSynchronize(map_.stream()));  // ds_ / src_ are complete before the other GPUs read them
    std::vector<Pose> poses((size_t)world);
    std::vector<int> iters((size_t)world, 0);
    std::vector<std::exception_ptr> errors((size_t)world);
    auto slice = [&](int r, size_t &lo, size_t &cnt) {
        lo = n_src_ * (size_t)r / (size_t)world;
        cnt = n_src_ * (size_t)(r + 1) / (size_t)world - lo;
    };
    auto work = [&](int r) {
        try {
            Replica &rep = *replicas_[(size_t)r - 1];
            SAGE_CUDA(cudaSetDevice(rep.device));
            size_t lo, cnt;
            slice(r, lo, cnt);
            rep.src.ensure(cnt ? cnt : 1);
            rep.ds.ensure(n_ds_ ? n_ds_ : 1);
            cudaStream_t st = rep.map->stream();
            if (cnt) SAGE_CUDA(cudaMemcpyPeerAsync(rep.src.p, rep.device, src_.p + lo, map_.device(), cnt * sizeof(double4), st));
            if (n_ds_) SAGE_CUDA(cudaMemcpyPeerAsync(rep.ds.p, rep.device, ds_.p, map_.device(), n_ds_ * sizeof(double4), st));
            iters[(size_t)r] = rep.map->register_frame_dev(rep.src.p, cnt, guess, max_dist, kernel, cfg_.sem_th, 500, 1e-4, poses[(size_t)r]);
            rep.map->update_dev(rep.ds.p, n_ds_, poses[(size_t)r]);
        } catch (...) {
            errors[(size_t)r] = std::current_exception();
        }
    };
    std::vector<std::thread> threads;
    for (int r = 1; r < world; ++r) threads.emplace_back(work, r);
    try {
        size_t lo, cnt;
        slice(0, lo, cnt);
        iters[0] = map_.register_frame_dev(src_.p, cnt, guess, max_dist, kernel, cfg_.sem_th, 500, 1e-4, poses[0]);
    } catch (...) {
        errors[0] = std::current_exception();
    }
    for (auto &t : threads) t.join();
    SAGE_CUDA(cudaSetDevice(map_.device()));
    for (auto &e : errors)
        if (e) std::rethrow_exception(e);
    for (int r = 1; r < world; ++r)
        if (std::memcmp(&poses[(size_t)r], &poses[0], sizeof(Pose)) != 0 || iters[(size_t)r] != iters[0])
            throw CudaError("sharded registration: the GPUs disagree on the pose (replicated maps out of step?)");
    pose_out = poses[0];
    return iters[0];
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
    if (replicas_.empty())
        last_iters_ = map_.register_frame_dev(src_.p, n_src_, initial_guess, 3.0 * sigma, sigma / 3.0, cfg_.sem_th, 500, 1e-4, new_pose);
    else
        last_iters_ = register_sharded(initial_guess, 3.0 * sigma, sigma / 3.0, new_pose);  // the replicas also update their maps
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
