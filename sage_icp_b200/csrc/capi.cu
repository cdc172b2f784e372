// extern "C" surface declared in include/sage_icp_b200.h.  Exceptions never cross the boundary: they become
// negative return codes + sage_last_error().
#include <cstring>
#include <memory>
#include <string>

#include "../../include/sage_icp_b200.h"
#include "nccl_shim.cuh"
#include "pipeline.cuh"

using namespace sage;

struct sage_map {
    VoxelMapGPU *impl;
    bool owned;
};
// deep copy of a sageConfig POD (the caller's arrays need not outlive sage_create)
struct ConfigCopy {
    sage_config_pod pod{};
    std::vector<int32_t> offsets, labels, basic, landmarks;
    std::vector<double> sizes;
    explicit ConfigCopy(const sage_config_pod &c) : pod(c) {
        if (c.n_groups < 0 || c.n_basic_parts_labels < 0 || c.n_dynamic_remove_lankmark < 0) throw ArgError("sageConfig: negative array length");
        if (c.n_groups > 0 && (!c.group_offsets || !c.voxel_size)) throw ArgError("sageConfig: voxel_size / group_offsets is NULL");
        if (c.n_groups > 0) offsets.assign(c.group_offsets, c.group_offsets + c.n_groups + 1), sizes.assign(c.voxel_size, c.voxel_size + c.n_groups);
        const int n_labels = offsets.empty() ? 0 : offsets.back();
        if (n_labels < 0 || (n_labels > 0 && !c.group_labels)) throw ArgError("sageConfig: group_labels is NULL or group_offsets negative");
        if (n_labels > 0) labels.assign(c.group_labels, c.group_labels + n_labels);
        if (c.n_basic_parts_labels > 0) {
            if (!c.basic_parts_labels) throw ArgError("sageConfig: basic_parts_labels is NULL");
            basic.assign(c.basic_parts_labels, c.basic_parts_labels + c.n_basic_parts_labels);
        }
        if (c.n_dynamic_remove_lankmark > 0) {
            if (!c.dynamic_remove_lankmark) throw ArgError("sageConfig: dynamic_remove_lankmark is NULL");
            landmarks.assign(c.dynamic_remove_lankmark, c.dynamic_remove_lankmark + c.n_dynamic_remove_lankmark);
        }
        pod.group_offsets = offsets.data(), pod.group_labels = labels.data(), pod.voxel_size = sizes.data();
        pod.basic_parts_labels = basic.data(), pod.dynamic_remove_lankmark = landmarks.data();
    }
    ConfigCopy(const ConfigCopy &) = delete;
    ConfigCopy &operator=(const ConfigCopy &) = delete;
};
struct sage_pipeline {
    Pipeline *impl;
    sage_map map_handle;
    std::vector<double> scratch;
    ConfigCopy *config;
    int device;
};

static thread_local std::string g_err;

template <class F>
static long long guarded(F &&f) {
    try {
        return (long long)f();
    } catch (const ArgError &e) {
        g_err = e.what();
        return SAGE_EINVAL;
    } catch (const CudaError &e) {
        g_err = e.what();
        return SAGE_ECUDA;
    } catch (const std::exception &e) {
        g_err = e.what();
        const bool nccl = std::strstr(e.what(), "nccl") || std::strstr(e.what(), "NCCL");
        return nccl ? SAGE_ENCCL : SAGE_ECUDA;
    }
}

// every entry point goes through these: a NULL handle is an error code, never a crash
static Pipeline &P(sage_pipeline *h) {
    if (!h || !h->impl) throw ArgError("null pipeline handle");
    return *h->impl;
}
static VoxelMapGPU &M(sage_map *m) {
    if (!m || !m->impl) throw ArgError("null map handle");
    return *m->impl;
}
static void need(const void *p, const char *what) {
    if (!p) throw ArgError(std::string("null argument: ") + what);
}

static long long copy_out(const std::vector<double> &v, double *out, size_t cap_points) {
    const size_t n = v.size() / 4;
    if (!out) return (long long)n;
    if (cap_points < n) {
        g_err = "output buffer too small";
        return SAGE_ECAPACITY;
    }
    if (n) std::memcpy(out, v.data(), v.size() * sizeof(double));
    return (long long)n;
}

extern "C" {

const char *sage_last_error(void) { return g_err.c_str(); }

int sage_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major == 10) ++ok;
    }
    return ok;
}

// ---- pipeline ---------------------------------------------------------------------------------
sage_pipeline *sage_create(const sage_config_pod *config, int device) {
    sage_pipeline *h = nullptr;
    const long long rc = guarded([&] {
        if (!config) throw ArgError("config is NULL");
        std::unique_ptr<ConfigCopy> copy(new ConfigCopy(*config));
        std::unique_ptr<Pipeline> p(new Pipeline(copy->pod, device));
        h = new sage_pipeline{p.get(), sage_map{&p->map(), false}, {}, copy.get(), device};
        p.release(), copy.release();
        return 0;
    });
    return rc == 0 ? h : nullptr;
}
void sage_destroy(sage_pipeline *h) {
    if (!h) return;
    delete h->impl;
    delete h->config;
    delete h;
}
int sage_set_devices(sage_pipeline *h, const int *ids, int n) {
    return (int)guarded([&] {
        P(h);
        need(ids, "ids");
        if (n < 1 || n > 8) throw ArgError("sage_set_devices: 1..8 GPUs of one node");
        if (!h->impl->poses().empty() || !h->impl->map().empty()) {
            if (n == 1 && ids[0] == h->device && h->impl->n_devices() == 1) return 0;
            throw ArgError("the devices can only be changed on a fresh or reinitialised pipeline");
        }
        if (ids[0] != h->device) {
            std::unique_ptr<Pipeline> p(new Pipeline(h->config->pod, ids[0]));  // throws if the device is unusable; the old one stays
            p->map().set_eviction_faithful(h->impl->map().eviction_faithful());
            delete h->impl;
            h->impl = p.release();
            h->map_handle = sage_map{&h->impl->map(), false};
            h->device = ids[0];
        }
        // n > 1: one replica of the map per extra GPU; RegisterFrame shards its ICP queries over all of them
        h->impl->set_replica_devices(std::vector<int>(ids + 1, ids + n));
        return 0;
    });
}
int sage_num_devices(sage_pipeline *h) {
    return (int)guarded([&] { return (long long)P(h).n_devices(); });
}
int sage_reset(sage_pipeline *h) {
    return (int)guarded([&] {
        P(h).reinitialize();
        return 0;
    });
}
int sage_register_frame(sage_pipeline *h, const double *xyzl, size_t n, const double *timestamps, double pose_out[7], double *t_icp,
                        double *t_all) {
    return (int)guarded([&] {
        if (!xyzl && n) throw ArgError("null argument: xyzl");
        need(pose_out, "pose_out");
        Pose p;
        double ti = 0, ta = 0;
        P(h).register_frame(xyzl, n, timestamps, p, ti, ta);
        pose_to_wire(p, pose_out);
        if (t_icp) *t_icp = ti;
        if (t_all) *t_all = ta;
        return 0;
    });
}
int sage_register_frame_pointcloud2(sage_pipeline *h, const uint8_t *data, size_t n_points, uint32_t point_step, uint32_t x_offset,
                                    uint32_t y_offset, uint32_t z_offset, uint32_t label_offset, int label_datatype, const double *timestamps,
                                    double pose_out[7], double *t_icp, double *t_all) {
    return (int)guarded([&] {
        if (!data && n_points) throw ArgError("null argument: data");
        need(pose_out, "pose_out");
        if (label_datatype != 2 && label_datatype != 7) throw ArgError("label_datatype must be 2 (UINT8) or 7 (FLOAT32)");
        Pose p;
        double ti = 0, ta = 0;
        P(h).register_frame_pointcloud2(data, n_points, point_step, x_offset, y_offset, z_offset, label_offset, label_datatype == 7,
                                            timestamps, p, ti, ta);
        pose_to_wire(p, pose_out);
        if (t_icp) *t_icp = ti;
        if (t_all) *t_all = ta;
        return 0;
    });
}
int64_t sage_last_source(sage_pipeline *h, double *out, size_t cap) {
    return guarded([&] {
        if (!out) return (long long)P(h).n_source();
        P(h).last_source(h->scratch);
        return copy_out(h->scratch, out, cap);
    });
}
int64_t sage_last_frame_downsample(sage_pipeline *h, double *out, size_t cap) {
    return guarded([&] {
        if (!out) return (long long)P(h).n_downsample();
        P(h).last_downsample(h->scratch);
        return copy_out(h->scratch, out, cap);
    });
}
int sage_last_iterations(sage_pipeline *h) {
    return (int)guarded([&] { return P(h).last_iterations(); });
}
double sage_last_sigma(sage_pipeline *h) {
    double v = -1.0;
    guarded([&] { return v = P(h).last_sigma(), 0; });
    return v;
}

int sage_voxelize(sage_pipeline *h, const double *xyzl, size_t n, double *source_out, size_t *n_source, double *downsample_out,
                  size_t *n_downsample) {
    return (int)guarded([&] {
        if (!xyzl && n) throw ArgError("null argument: xyzl");
        need(n_source, "n_source"), need(n_downsample, "n_downsample");
        std::vector<double> s, d;
        P(h).voxelize_host(xyzl, n, s, d);
        *n_source = s.size() / 4, *n_downsample = d.size() / 4;
        if (source_out && !s.empty()) std::memcpy(source_out, s.data(), s.size() * sizeof(double));
        if (downsample_out && !d.empty()) std::memcpy(downsample_out, d.data(), d.size() * sizeof(double));
        return 0;
    });
}
double sage_get_adaptive_threshold(sage_pipeline *h) {
    double v = -1.0;
    guarded([&] { return v = P(h).get_adaptive_threshold(), 0; });
    return v;
}
int sage_has_moved(sage_pipeline *h) {
    return (int)guarded([&] { return P(h).has_moved() ? 1 : 0; });
}
int sage_get_prediction_model(sage_pipeline *h, double pose_out[7]) {
    return (int)guarded([&] {
        need(pose_out, "pose_out");
        pose_to_wire(P(h).get_prediction_model(), pose_out);
        return 0;
    });
}
int sage_transform_to_last_frame(sage_pipeline *h, const double last_pose[7], const double current_pose[7], const double *xyzl, size_t n,
                                 double *out) {
    (void)h;
    if (!last_pose || !current_pose || (n && (!xyzl || !out))) {
        g_err = "null argument";
        return SAGE_EINVAL;
    }
    // TransformPoints(last_pose.inverse() * current_pose, points): a few thousand points for RViz; host arithmetic
    const Pose T = pose_mul(pose_inverse(pose_from_wire(last_pose)), pose_from_wire(current_pose));
    for (size_t i = 0; i < n; ++i) {
        double x, y, z;
        pose_act(T, xyzl[4 * i], xyzl[4 * i + 1], xyzl[4 * i + 2], x, y, z);
        const double l = xyzl[4 * i + 3];
        out[4 * i] = x, out[4 * i + 1] = y, out[4 * i + 2] = z, out[4 * i + 3] = l;
    }
    return 0;
}
int sage_key_frame_grid(sage_pipeline *h, const double *xyzl, size_t n, const double *last_pose, const double *current_pose,
                        const double bounds[6], int rows, int cols, const int32_t *last_occ, int32_t *grid_out, double *overlap_out) {
    return (int)guarded([&] {
        need(bounds, "bounds");
        if (!xyzl && n) throw ArgError("null argument: xyzl");
        if ((last_pose == nullptr) != (current_pose == nullptr)) throw ArgError("last_pose and current_pose go together");
        if (last_occ && !overlap_out) throw ArgError("null argument: overlap_out");
        Pose a, b;
        if (last_pose) a = pose_from_wire(last_pose), b = pose_from_wire(current_pose);
        P(h).key_frame_grid(xyzl, n, last_pose ? &a : nullptr, last_pose ? &b : nullptr, bounds, rows, cols, last_occ, grid_out, overlap_out);
        return 0;
    });
}
int64_t sage_num_poses(sage_pipeline *h) {
    return guarded([&] { return (long long)P(h).poses().size(); });
}
int sage_get_pose(sage_pipeline *h, size_t i, double pose_out[7]) {
    return (int)guarded([&] {
        need(pose_out, "pose_out");
        if (i >= P(h).poses().size()) throw ArgError("pose index out of range");
        pose_to_wire(P(h).poses()[i], pose_out);
        return 0;
    });
}
int64_t sage_get_poses(sage_pipeline *h, size_t first, double *poses_out, size_t cap) {
    return guarded([&] {
        const auto &ps = P(h).poses();
        if (first > ps.size()) throw ArgError("pose index out of range");
        const size_t avail = ps.size() - first;
        if (!poses_out || cap == 0) return (long long)avail;  // query: how many poses from `first` on
        const size_t n = avail < cap ? avail : cap;
        for (size_t i = 0; i < n; ++i) pose_to_wire(ps[first + i], poses_out + 7 * i);
        return (long long)n;
    });
}
int64_t sage_local_map(sage_pipeline *h, double *out, size_t cap) {
    return guarded([&] { return P(h).map().pointcloud(out, cap); });
}
sage_map *sage_pipeline_map(sage_pipeline *h) { return h ? &h->map_handle : nullptr; }
int64_t sage_preprocess(sage_pipeline *h, const double *xyzl, size_t n, double *out, size_t cap) {
    return guarded([&] {
        if (!xyzl && n) throw ArgError("null argument: xyzl");
        P(h).preprocess_host(xyzl, n, h->scratch);
        return copy_out(h->scratch, out, cap);
    });
}
int64_t sage_voxel_downsample(sage_pipeline *h, const double *xyzl, size_t n, double vox_scale, double *out, size_t cap) {
    return guarded([&] {
        if (!xyzl && n) throw ArgError("null argument: xyzl");
        P(h).downsample_host(xyzl, n, vox_scale, h->scratch);
        return copy_out(h->scratch, out, cap);
    });
}

// ---- map --------------------------------------------------------------------------------------
sage_map *sage_map_create(double voxel_size, double max_distance, int basic, int critical, const int32_t *labels, int n_labels, int device) {
    sage_map *m = nullptr;
    const long long rc = guarded([&] {
        m = new sage_map{new VoxelMapGPU(voxel_size, max_distance, basic, critical, labels, n_labels, device), true};
        return 0;
    });
    return rc == 0 ? m : nullptr;
}
void sage_map_destroy(sage_map *m) {
    if (!m || !m->owned) return;
    delete m->impl;
    delete m;
}
int sage_map_clear(sage_map *m) {
    return (int)guarded([&] {
        M(m).clear();
        return 0;
    });
}
int sage_map_set_eviction(sage_map *m, int faithful) {
    return (int)guarded([&] {
        if (!m) throw ArgError("null argument");
        M(m).set_eviction_faithful(faithful != 0);
        return 0;
    });
}
int sage_map_empty(sage_map *m) {
    return (int)guarded([&] { return M(m).empty() ? 1 : 0; });
}
int64_t sage_map_num_voxels(sage_map *m) {
    return guarded([&] { return M(m).num_voxels(); });
}
int64_t sage_map_num_points(sage_map *m) {
    return guarded([&] { return M(m).num_points(); });
}
int sage_map_add_points(sage_map *m, const double *xyzl, size_t n) {
    return (int)guarded([&] {
        if (!xyzl && n) throw ArgError("null argument: xyzl");
        M(m).add_points_host(xyzl, n, nullptr);
        return 0;
    });
}
int sage_map_remove_far(sage_map *m, const double origin[3]) {
    return (int)guarded([&] {
        need(origin, "origin");
        M(m).remove_far(origin[0], origin[1], origin[2]);
        return 0;
    });
}
int sage_map_update(sage_map *m, const double *xyzl, size_t n, const double pose[7]) {
    return (int)guarded([&] {
        need(pose, "pose");
        if (!xyzl && n) throw ArgError("null argument: xyzl");
        const Pose T = pose_from_wire(pose);
        M(m).add_points_host(xyzl, n, &T);
        M(m).remove_far(T.tx, T.ty, T.tz);
        return 0;
    });
}
int64_t sage_map_pointcloud(sage_map *m, double *out, size_t cap) {
    return guarded([&] { return M(m).pointcloud(out, cap); });
}
int sage_map_load(sage_map *m, const int32_t *keys, const int32_t *counts, const double *points, int stride, size_t n_voxels) {
    return (int)guarded([&] {
        if (n_voxels) need(keys, "keys"), need(counts, "counts"), need(points, "points");
        M(m).load(keys, counts, points, stride, n_voxels);
        return 0;
    });
}
int64_t sage_map_dump(sage_map *m, int32_t *keys, int32_t *counts, double *points, size_t cap_voxels) {
    return guarded([&] { return M(m).dump(keys, counts, points, cap_voxels); });
}
int64_t sage_map_get_correspondences(sage_map *m, const double *xyzl, size_t n, double max_dist, double th, double *target_out,
                                     uint8_t *matched_out) {
    return guarded([&] {
        if (n) need(xyzl, "xyzl"), need(target_out, "target_out"), need(matched_out, "matched_out");
        return M(m).get_correspondences(xyzl, n, max_dist, th, target_out, matched_out);
    });
}
int sage_map_nn_stats(sage_map *m, const double *xyzl, size_t n, uint64_t *occupied, uint64_t *candidates) {
    return (int)guarded([&] {
        need(occupied, "occupied"), need(candidates, "candidates");
        if (!xyzl && n) throw ArgError("null argument: xyzl");
        unsigned long long o = 0, c = 0;
        M(m).nn_stats(xyzl, n, &o, &c);
        *occupied = o, *candidates = c;
        return 0;
    });
}
int sage_map_search_work(sage_map *m, const double *xyzl, size_t n, double max_dist, double th, uint64_t *scanned, uint64_t *probes,
                         uint64_t *exact, uint64_t *deferred, uint64_t *staged) {
    return (int)guarded([&] {
        need(scanned, "scanned"), need(probes, "probes"), need(exact, "exact");
        if (!xyzl && n) throw ArgError("null argument: xyzl");
        unsigned long long a = 0, b = 0, c = 0, d = 0, e = 0;
        M(m).search_work(xyzl, n, max_dist, th, &a, &b, &c, &d, &e);
        *scanned = a, *probes = b, *exact = c;
        if (deferred) *deferred = d;
        if (staged) *staged = e;
        return 0;
    });
}
int sage_core_register_frame(sage_map *m, const double *frame, size_t n, const double guess[7], double max_dist, double kernel,
                             double sem_th, int max_iters, double est_th, double pose_out[7], int *iters_out) {
    return (int)guarded([&] {
        need(guess, "guess"), need(pose_out, "pose_out");
        if (!frame && n) throw ArgError("null argument: frame");
        Pose out;
        const int it = M(m).register_frame_host(frame, n, pose_from_wire(guess), max_dist, kernel, sem_th, max_iters, est_th, out);
        pose_to_wire(out, pose_out);
        if (iters_out) *iters_out = it;
        return 0;
    });
}
int sage_core_register_frame_device(sage_map *m, const void *frame_dev, size_t n, const double guess[7], double max_dist, double kernel,
                                    double sem_th, int max_iters, double est_th, double pose_out[7], int *iters_out) {
    return (int)guarded([&] {
        need(guess, "guess"), need(pose_out, "pose_out");
        if (!frame_dev && n) throw ArgError("null argument: frame_dev");
        Pose out;
        const int it = M(m).register_frame_dev((const double4 *)frame_dev, n, pose_from_wire(guess), max_dist, kernel, sem_th,
                                                   max_iters, est_th, out);
        pose_to_wire(out, pose_out);
        if (iters_out) *iters_out = it;
        return 0;
    });
}
int sage_core_normal_equations(sage_map *m, const double *frame, size_t n, double max_dist, double kernel, double sem_th,
                               double JTJ[36], double JTr[6], int64_t *pairs) {
    return (int)guarded([&] {
        need(JTJ, "JTJ"), need(JTr, "JTr");
        if (!frame && n) throw ArgError("null argument: frame");
        long long np = 0;
        M(m).normal_equations(frame, n, max_dist, kernel, sem_th, JTJ, JTr, &np);
        if (pairs) *pairs = np;
        return 0;
    });
}

// ---- measurement / multi-GPU ------------------------------------------------------------------
void *sage_map_stream(sage_map *m) { return m && m->impl ? (void *)m->impl->stream() : nullptr; }
int sage_map_profile_enable(sage_map *m, int enable) {
    return (int)guarded([&] {
        M(m).profile_enable(enable != 0);
        return 0;
    });
}
int sage_map_profile_read(sage_map *m, int64_t *iterations, double *total_ms) {
    return (int)guarded([&] {
        long long l = 0, k = 0;
        double ms = 0;
        M(m).profile_read(&l, &ms, &k);
        if (iterations) *iterations = l;
        if (total_ms) *total_ms = ms;
        return 0;
    });
}
int sage_map_profile_read_launches(sage_map *m, int64_t *iterations, int64_t *launches, double *total_ms) {
    return (int)guarded([&] {
        long long l = 0, k = 0;
        double ms = 0;
        M(m).profile_read(&l, &ms, &k);
        if (iterations) *iterations = l;
        if (launches) *launches = k;
        if (total_ms) *total_ms = ms;
        return 0;
    });
}
// development aid, deliberately not in the public header
size_t sage_debug_timeline(sage_map *m, unsigned long long *out, size_t cap) {
    const long long rc = guarded([&] { return (long long)M(m).debug_timeline(out, cap); });
    return rc < 0 ? 0 : (size_t)rc;
}
int64_t sage_launch_count(void) { return g_launches.load(); }

int sage_robin_iteration_order(const uint32_t *hash20, size_t n, uint32_t *order_out) {
    return (int)guarded([&] {
        if (n && (!hash20 || !order_out)) throw ArgError("null argument");
        robin_iteration_order(hash20, n, order_out);
        return 0;
    });
}
int64_t sage_robin_table_replay(const int32_t *ops, size_t n_ops, int32_t *keys_out, size_t cap, uint64_t *bucket_count) {
    return guarded([&]() -> int64_t {
        if ((n_ops && !ops) || (cap && !keys_out)) throw ArgError("null argument");
        HostVoxelTable t;
        for (size_t i = 0; i < n_ops; ++i) {
            const int32_t *o = ops + 5 * i;
            if (o[0] == 0) {
                if (!key_in_range(o[1], o[2], o[3])) throw ArgError("voxel key outside packable range");
                const unsigned long long key = pack_key(o[1], o[2], o[3]);
                t.insert(key, reference_voxel_hash(key), (uint32_t)i);
            } else if (o[0] == 1) {
                for (size_t b = 0; b < t.bucket_count(); ++b) {
                    if (t.at(b).dist < 0) continue;
                    int x, y, z;
                    unpack_key(t.at(b).key, x, y, z);
                    const long long dx = (long long)x - o[1], dy = (long long)y - o[2], dz = (long long)z - o[3];
                    if (dx * dx + dy * dy + dz * dz > (long long)o[4]) t.erase_at(b);
                }
            } else if (o[0] == 2) {
                t.clear();
            } else {
                throw ArgError("unknown op");
            }
        }
        size_t k = 0;
        for (size_t b = 0; b < t.bucket_count(); ++b) {
            if (t.at(b).dist < 0) continue;
            if (k < cap) unpack_key(t.at(b).key, keys_out[3 * k], keys_out[3 * k + 1], keys_out[3 * k + 2]);
            ++k;
        }
        if (bucket_count) *bucket_count = t.bucket_count();
        return (int64_t)k;
    });
}
int sage_shard_range(size_t n, int rank, int world, size_t *begin, size_t *end) {
    if (world < 1 || rank < 0 || rank >= world || !begin || !end) {
        g_err = "bad shard arguments";
        return SAGE_EINVAL;
    }
    // contiguous, balanced to within one query: rank r owns [n*r/world, n*(r+1)/world)
    *begin = (size_t)(((unsigned __int128)n * (unsigned)rank) / (unsigned)world);
    *end = (size_t)(((unsigned __int128)n * (unsigned)(rank + 1)) / (unsigned)world);
    return 0;
}
int sage_nccl_unique_id(uint8_t id_out[128]) {
    return (int)guarded([&] {
        need(id_out, "id_out");
        nccl_unique_id(id_out);
        return 0;
    });
}
int sage_map_comm_init(sage_map *m, int rank, int world, const uint8_t id[128]) {
    return (int)guarded([&] {
        need(id, "id");
        if (world < 1 || rank < 0 || rank >= world) throw ArgError("bad rank / world");
        M(m).comm_init(rank, world, id);
        return 0;
    });
}
int sage_map_comm_peer_handle(sage_map *m, uint8_t handle_out[64]) {
    return (int)guarded([&] {
        need(handle_out, "handle_out");
        M(m).peer_handle(handle_out);
        return 0;
    });
}
int sage_map_comm_peer_attach(sage_map *m, int rank, int world, const uint8_t *handles) {
    return (int)guarded([&] {
        need(handles, "handles");
        M(m).peer_attach(rank, world, handles);
        return 0;
    });
}
int sage_map_comm_destroy(sage_map *m) {
    return (int)guarded([&] {
        M(m).comm_destroy();
        M(m).peer_detach();
        return 0;
    });
}

}  // extern "C"
