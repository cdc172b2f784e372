// ORACLE — TEST INFRASTRUCTURE ONLY. C entry points over sage_oracle.hpp so that tests/, smoke() and
// bench.py's cpu_baseline / --impl reference legs can drive the CPU restatement through ctypes.
// Pose wire format everywhere: double[7] = tx, ty, tz, qx, qy, qz, qw.
#include <cstdint>
#include <cstring>
#include <vector>

#include "sage_oracle.hpp"

using namespace orc;

extern "C" {

// Same field order/meaning as include/sage_icp_b200.h:sage_config_pod (kept layout-identical on purpose so
// one ctypes.Structure serves both; the two headers do not include each other).
struct orc_config_pod {
    int32_t n_groups;
    const int32_t *group_offsets;  // n_groups+1 offsets into group_labels
    const int32_t *group_labels;
    const double *voxel_size;  // n_groups
    double voxel_size_map, max_range, min_range, label_max_range, local_map_range;
    int32_t basic_points_per_voxel, critical_points_per_voxel;
    int32_t n_basic_parts_labels;
    const int32_t *basic_parts_labels;
    double min_motion_th, initial_threshold, sem_th;
    int32_t deskew, dynamic_vehicle_filter;
    double dynamic_vehicle_filter_th;
    int32_t dynamic_vehicle_voxid;
    int32_t n_dynamic_remove_lankmark;
    const int32_t *dynamic_remove_lankmark;
};
}

static Config to_config(const orc_config_pod *p) {
    Config c;
    for (int g = 0; g < p->n_groups; ++g) {
        c.voxel_labels.emplace_back(p->group_labels + p->group_offsets[g], p->group_labels + p->group_offsets[g + 1]);
        c.voxel_size.push_back(p->voxel_size[g]);
    }
    c.voxel_size_map = p->voxel_size_map;
    c.max_range = p->max_range;
    c.min_range = p->min_range;
    c.label_max_range = p->label_max_range;
    c.local_map_range = p->local_map_range;
    c.basic_points_per_voxel = p->basic_points_per_voxel;
    c.critical_points_per_voxel = p->critical_points_per_voxel;
    c.basic_parts_labels.assign(p->basic_parts_labels, p->basic_parts_labels + p->n_basic_parts_labels);
    c.min_motion_th = p->min_motion_th;
    c.initial_threshold = p->initial_threshold;
    c.sem_th = p->sem_th;
    c.deskew = p->deskew != 0;
    c.dynamic_vehicle_filter = p->dynamic_vehicle_filter != 0;
    c.dynamic_vehicle_filter_th = p->dynamic_vehicle_filter_th;
    c.dynamic_vehicle_voxid = p->dynamic_vehicle_voxid;
    if (p->n_dynamic_remove_lankmark > 0)
        c.dynamic_remove_lankmark.assign(p->dynamic_remove_lankmark, p->dynamic_remove_lankmark + p->n_dynamic_remove_lankmark);
    return c;
}

static Cloud to_cloud(const double *xyzl, size_t n) {
    Cloud c(n);
    if (n) std::memcpy(c.data(), xyzl, n * sizeof(Point4));
    return c;
}
static size_t from_cloud(const Cloud &c, double *out, size_t cap) {
    if (out && cap >= c.size() && !c.empty()) std::memcpy(out, c.data(), c.size() * sizeof(Point4));
    return c.size();
}
static SE3 to_se3(const double p[7]) {
    SE3 T;
    T.t = {p[0], p[1], p[2]};
    T.q = {p[6], p[3], p[4], p[5]};
    return T;
}
static void from_se3(const SE3 &T, double p[7]) {
    p[0] = T.t.x, p[1] = T.t.y, p[2] = T.t.z, p[3] = T.q.x, p[4] = T.q.y, p[5] = T.q.z, p[6] = T.q.w;
}

struct OrcPipeline {
    SageICP icp;
    Cloud last_source;
    explicit OrcPipeline(const Config &c) : icp(c) {}
};

extern "C" {

// ---- SE(3) helpers -------------------------------------------------------------------------
void orc_se3_exp(const double xi[6], double pose[7]) { from_se3(se3_exp(xi), pose); }
void orc_se3_log(const double pose[7], double xi[6]) { se3_log(to_se3(pose), xi); }
void orc_se3_mul(const double a[7], const double b[7], double out[7]) { from_se3(se3_mul(to_se3(a), to_se3(b)), out); }
void orc_se3_inverse(const double a[7], double out[7]) { from_se3(se3_inverse(to_se3(a)), out); }
void orc_se3_act(const double a[7], const double p[3], double out[3]) {
    const Vec3 r = se3_act(to_se3(a), {p[0], p[1], p[2]});
    out[0] = r.x, out[1] = r.y, out[2] = r.z;
}
double orc_rotation_angle(const double pose[7]) { return angle_of_rotation_matrix(quat_matrix(to_se3(pose).q)); }
void orc_ldlt6_solve(const double A[36], const double b[6], double x[6]) {
    double M[6][6];
    std::memcpy(M, A, sizeof(M));
    ldlt6_solve(M, b, x);
}
uint32_t orc_voxel_hash(int32_t x, int32_t y, int32_t z) { return voxel_hash({x, y, z}); }

// robin-table iteration-order probe: insert distinct keys in order, write the iteration order (indices into
// the input) to order_out; returns the final bucket_count.
size_t orc_robin_order(const int32_t *keys, size_t n, int64_t *order_out) {
    RobinTable<int64_t> t;
    for (size_t i = 0; i < n; ++i) t.insert({keys[3 * i], keys[3 * i + 1], keys[3 * i + 2]}, (int64_t)i);
    size_t k = 0;
    for (const auto &b : t.buckets())
        if (!b.empty()) order_out[k++] = b.value;
    return t.bucket_count();
}

// General replay for the third-party self-check (tools/verify_thirdparty.cpp): ops[i] = {kind, x, y, z}; kind 0 = insert(key) if
// absent, 1 = erase(key) if present, 2 = the reference's erase-while-iterating sweep (core/VoxelHashMap.cpp:176-184) erasing
// every key with x < ops[i].x that the iterator meets.  Writes the final iteration order (the id each key got at its insertion:
// its op index); returns its length.
size_t orc_robin_replay(const int32_t *ops, size_t n_ops, int64_t *order_out, int64_t *bucket_count_out) {
    RobinTable<int64_t> t;
    for (size_t i = 0; i < n_ops; ++i) {
        const int32_t *o = ops + 4 * i;
        const Voxel k{o[1], o[2], o[3]};
        if (o[0] == 0) {
            if (!t.contains(k)) t.insert(k, (int64_t)i);
        } else if (o[0] == 1) {
            const size_t ib = t.find(k);
            if (ib != t.npos) t.erase_at(ib);
        } else {
            auto &B = t.buckets();
            for (size_t ib = 0; ib < B.size(); ++ib)
                if (!B[ib].empty() && B[ib].key.x < o[1]) t.erase_at(ib);
        }
    }
    size_t k = 0;
    for (const auto &b : t.buckets())
        if (!b.empty()) order_out[k++] = b.value;
    if (bucket_count_out) *bucket_count_out = (int64_t)t.bucket_count();
    return k;
}

// ---- core free functions -------------------------------------------------------------------
size_t orc_preprocess(const double *xyzl, size_t n, double max_range, double min_range, double label_max_range,
                      double *out, size_t cap) {
    return from_cloud(Preprocess(to_cloud(xyzl, n), max_range, min_range, label_max_range), out, cap);
}
size_t orc_preprocess_dynamic_ordered(const orc_config_pod *cfg, const double *xyzl, size_t n, int cluster_order, double *out, size_t cap) {
    const Config c = to_config(cfg);
    return from_cloud(PreprocessDynamic(to_cloud(xyzl, n), c.max_range, c.min_range, c.label_max_range, c.dynamic_vehicle_filter_th,
                                        c.voxel_labels[(size_t)c.dynamic_vehicle_voxid], c.dynamic_remove_lankmark, cluster_order != 0),
                      out, cap);
}
size_t orc_preprocess_dynamic(const orc_config_pod *cfg, const double *xyzl, size_t n, double *out, size_t cap) {
    return orc_preprocess_dynamic_ordered(cfg, xyzl, n, 0, out, cap);
}
size_t orc_voxel_downsample(const orc_config_pod *cfg, const double *xyzl, size_t n, double vox_scale, double *out, size_t cap) {
    const Config c = to_config(cfg);
    return from_cloud(VoxelDownsample(to_cloud(xyzl, n), c.voxel_labels, c.voxel_size, vox_scale), out, cap);
}

// ---- VoxelHashMap ---------------------------------------------------------------------------
void *orc_map_create(double voxel_size, double max_distance, int basic, int critical, const int32_t *labels, int n_labels,
                     int evict_faithful) {
    auto *m = new VoxelHashMap(voxel_size, max_distance, basic, critical, std::vector<int>(labels, labels + n_labels));
    m->evict_faithful_ = evict_faithful != 0;
    return m;
}
void orc_map_destroy(void *m) { delete (VoxelHashMap *)m; }
void orc_map_clear(void *m) { ((VoxelHashMap *)m)->Clear(); }
size_t orc_map_num_voxels(void *m) { return ((VoxelHashMap *)m)->map_.size(); }
size_t orc_map_bucket_count(void *m) { return ((VoxelHashMap *)m)->map_.bucket_count(); }
size_t orc_map_num_points(void *m) {
    size_t n = 0;
    for (const auto &b : ((VoxelHashMap *)m)->map_.buckets())
        if (!b.empty()) n += b.value.points.size();
    return n;
}
void orc_map_add_points(void *m, const double *xyzl, size_t n) { ((VoxelHashMap *)m)->AddPoints(to_cloud(xyzl, n)); }
void orc_map_remove_far(void *m, const double origin[3]) {
    ((VoxelHashMap *)m)->RemovePointsFarFromLocation({origin[0], origin[1], origin[2]});
}
void orc_map_update(void *m, const double *xyzl, size_t n, const double pose[7]) {
    ((VoxelHashMap *)m)->Update(to_cloud(xyzl, n), to_se3(pose));
}
size_t orc_map_pointcloud(void *m, double *out, size_t cap) { return from_cloud(((VoxelHashMap *)m)->Pointcloud(), out, cap); }
// Dump in iteration order: keys V x 3, counts V, points V x stride x 4 (unused tail zero). Returns V.
size_t orc_map_dump(void *m, int32_t *keys, int32_t *counts, double *pts, int stride, size_t cap_voxels) {
    auto *M = (VoxelHashMap *)m;
    if (!keys || cap_voxels < M->map_.size()) return M->map_.size();
    size_t v = 0;
    for (const auto &b : M->map_.buckets()) {
        if (b.empty()) continue;
        keys[3 * v] = b.key.x, keys[3 * v + 1] = b.key.y, keys[3 * v + 2] = b.key.z;
        counts[v] = (int32_t)b.value.points.size();
        std::memset(pts + v * (size_t)stride * 4, 0, sizeof(double) * 4 * (size_t)stride);
        std::memcpy(pts + v * (size_t)stride * 4, b.value.points.data(), sizeof(Point4) * b.value.points.size());
        ++v;
    }
    return v;
}
// GetCorrespondences; outputs sized n. qidx (optional) = query index of each pair. Returns number of pairs.
size_t orc_map_get_correspondences(void *m, const double *xyzl, size_t n, double max_dist, double th, int threads,
                                   double *src_out, double *tgt_out, int64_t *qidx) {
    const auto c = ((VoxelHashMap *)m)->GetCorrespondences(to_cloud(xyzl, n), max_dist, th, threads);
    from_cloud(c.source, src_out, n);
    from_cloud(c.target, tgt_out, n);
    if (qidx) std::memcpy(qidx, c.query_index.data(), c.query_index.size() * sizeof(int64_t));
    return c.source.size();
}
// key-frame occupancy grid / overlap of the ROS node (ros/ros2/Utils.hpp:220-258); bounds6 = {x0,x1,y0,y1,z0,z1}
void orc_grid_map(const double *xyzl, size_t n, const double *bounds6, int H, int W, int32_t *grid_out) {
    const double b[3][2] = {{bounds6[0], bounds6[1]}, {bounds6[2], bounds6[3]}, {bounds6[4], bounds6[5]}};
    const auto g = EigenToGridMap(to_cloud(xyzl, n), b, H, W);
    for (size_t i = 0; i < g.size(); ++i) grid_out[i] = g[i];
}
double orc_occ_overlap(const int32_t *occ_s, const int32_t *occ_t, size_t cells) {
    return ComputeOccOverlap(std::vector<int>(occ_s, occ_s + cells), std::vector<int>(occ_t, occ_t + cells));
}
// exact occupancy statistics for the algorithmic-bytes formula: sum over queries of occupied
// neighbour voxels and of candidate points.
void orc_map_nn_stats(void *m, const double *xyzl, size_t n, uint64_t *occupied, uint64_t *candidates) {
    auto *M = (VoxelHashMap *)m;
    uint64_t o = 0, c = 0;
    for (size_t q = 0; q < n; ++q) {
        const int kx = (int)(xyzl[4 * q] / M->voxel_size_), ky = (int)(xyzl[4 * q + 1] / M->voxel_size_),
                  kz = (int)(xyzl[4 * q + 2] / M->voxel_size_);
        for (int i = kx - 1; i <= kx + 1; ++i)
            for (int j = ky - 1; j <= ky + 1; ++j)
                for (int k = kz - 1; k <= kz + 1; ++k) {
                    const size_t ib = M->map_.find({i, j, k});
                    if (ib != M->map_.npos) ++o, c += M->map_.value_at(ib).points.size();
                }
    }
    *occupied = o, *candidates = c;
}

// ---- Registration ---------------------------------------------------------------------------
// JTJ row-major 36, JTr 6, x 6 (solution of JTJ x = -JTr), est pose 7
void orc_align_clouds(const double *src, const double *tgt, size_t n, double th, int threads, double *JTJ, double *JTr,
                      double *x, double *est) {
    NormalEq ne;
    const SE3 e = AlignClouds(to_cloud(src, n), to_cloud(tgt, n), th, threads, &ne);
    if (JTJ) std::memcpy(JTJ, ne.JTJ, sizeof(ne.JTJ));
    if (JTr) std::memcpy(JTr, ne.JTr, sizeof(ne.JTr));
    if (x) {
        double nb[6];
        for (int i = 0; i < 6; ++i) nb[i] = -ne.JTr[i];
        ldlt6_solve(ne.JTJ, nb, x);
    }
    if (est) from_se3(e, est);
}
int orc_register_frame_core(void *m, const double *frame, size_t n, const double guess[7], double max_dist, double kernel,
                            double sem_th, int threads, int max_iters, double est_th, double pose_out[7]) {
    int iters = 0;
    const SE3 T = RegisterFrameCore(to_cloud(frame, n), *(VoxelHashMap *)m, to_se3(guess), max_dist, kernel, sem_th, threads,
                                    max_iters, est_th, &iters);
    from_se3(T, pose_out);
    return iters;
}
size_t orc_deskew(const double *frame, const double *ts, size_t n, const double start[7], const double finish[7], double *out) {
    return from_cloud(DeSkewScan(to_cloud(frame, n), std::vector<double>(ts, ts + n), to_se3(start), to_se3(finish)), out, n);
}

// ---- pipeline (sageICP) ---------------------------------------------------------------------
void *orc_create(const orc_config_pod *cfg, int threads, int evict_faithful) {
    auto *p = new OrcPipeline(to_config(cfg));
    p->icp.threads_ = threads;
    p->icp.sem_map_.evict_faithful_ = evict_faithful != 0;
    return p;
}
void orc_destroy(void *h) { delete (OrcPipeline *)h; }
void orc_reset(void *h) { ((OrcPipeline *)h)->icp.reinitialize(); }
void orc_set_dynamic_cluster_order(void *h, int on) { ((OrcPipeline *)h)->icp.dynamic_cluster_order_ = on != 0; }
int orc_register_frame(void *h, const double *xyzl, size_t n, const double *ts, double pose_out[7], double *t_icp, double *t_all) {
    auto *p = (OrcPipeline *)h;
    const Cloud frame = to_cloud(xyzl, n);
    auto res = ts ? p->icp.RegisterFrame(frame, std::vector<double>(ts, ts + n)) : p->icp.RegisterFrame(frame);
    p->last_source = std::move(std::get<0>(res));
    if (t_icp) *t_icp = std::get<1>(res);
    if (t_all) *t_all = std::get<2>(res);
    from_se3(p->icp.poses_.back(), pose_out);
    return 0;
}
size_t orc_voxelize(void *h, const double *xyzl, size_t n, double *source_out, size_t *n_source, double *ds_out, size_t *n_ds) {
    auto [source, ds] = ((OrcPipeline *)h)->icp.Voxelize(to_cloud(xyzl, n));
    *n_source = from_cloud(source, source_out, n);
    *n_ds = from_cloud(ds, ds_out, n);
    return *n_source;
}
double orc_get_adaptive_threshold(void *h) { return ((OrcPipeline *)h)->icp.GetAdaptiveThreshold(); }
int orc_has_moved(void *h) { return ((OrcPipeline *)h)->icp.HasMoved() ? 1 : 0; }
void orc_get_prediction_model(void *h, double out[7]) { from_se3(((OrcPipeline *)h)->icp.GetPredictionModel(), out); }
size_t orc_last_source(void *h, double *out, size_t cap) { return from_cloud(((OrcPipeline *)h)->last_source, out, cap); }
size_t orc_last_frame_downsample(void *h, double *out, size_t cap) {
    return from_cloud(((OrcPipeline *)h)->icp.last_frame_downsample_, out, cap);
}
int orc_last_iterations(void *h) { return ((OrcPipeline *)h)->icp.last_iterations_; }
double orc_last_sigma(void *h) { return ((OrcPipeline *)h)->icp.last_sigma_; }
size_t orc_num_poses(void *h) { return ((OrcPipeline *)h)->icp.poses_.size(); }
void orc_get_pose(void *h, size_t i, double out[7]) { from_se3(((OrcPipeline *)h)->icp.poses_[i], out); }
size_t orc_local_map(void *h, double *out, size_t cap) { return from_cloud(((OrcPipeline *)h)->icp.LocalMap(), out, cap); }
void *orc_pipeline_map(void *h) { return &((OrcPipeline *)h)->icp.sem_map_; }
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
