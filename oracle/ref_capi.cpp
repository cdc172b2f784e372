// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// C entry points over the REFERENCE's own classes and free functions, compiled from the reference's unmodified sources where
// they lie (/root/reference/cpp/sage_icp/core/{Deskew,Preprocessing,Registration,Threshold,VoxelHashMap}.cpp and
// pipeline/sageICP.cpp) against the stand-in headers of oracle/shim/ — Eigen, Sophus, oneTBB, tsl::robin_map and PCL are not in
// this image.  The result, oracle/_ref/libsage_ref.so (oracle/Makefile, target `ref`), is used by tests/test_reference_build.py
// to check the oracle's restatement against the reference's OWN code on the same seeded inputs: loop structure, rule tables,
// thresholds, call order, the semantic metric, the AddPoint table, the eviction sweep, the pipeline's state machine.  What it
// cannot pin is the third-party arithmetic itself (pivoted LDLT, SE3 exp/log, robin_map bucket order, FLANN/PCL order): the
// stand-ins implement those from the same published formulas the oracle uses.
// No reference source is copied into this repository; this file only calls the reference's public interface:
//   sage_icp::VoxelHashMap            core/VoxelHashMap.hpp:35-107
//   sage_icp::RegisterFrame           core/Registration.hpp:34-39
//   sage_icp::Preprocess / VoxelDownsample   core/Preprocessing.hpp:35-48
//   sage_icp::DeSkewScan              core/Deskew.hpp:31-34
//   sage_icp::pipeline::sageICP       pipeline/sageICP.hpp:67-109
#include <cstdint>
#include <cstring>
#include <vector>

#include "sage_icp/core/Deskew.hpp"
#include "sage_icp/core/Preprocessing.hpp"
#include "sage_icp/core/Registration.hpp"
#include "sage_icp/core/VoxelHashMap.hpp"
#include "sage_icp/pipeline/sageICP.hpp"

using Cloud = std::vector<Eigen::Vector4d>;

namespace {

struct ref_config_pod {  // same layout as sage_config_pod (include/sage_icp_b200.h) and orc_config_pod
    int32_t n_groups;
    const int32_t *group_offsets;
    const int32_t *group_labels;
    const double *voxel_size;
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

sage_icp::pipeline::sageConfig to_config(const ref_config_pod *p) {
    sage_icp::pipeline::sageConfig c;
    for (int g = 0; g < p->n_groups; ++g) {
        c.voxel_labels.emplace_back(p->group_labels + p->group_offsets[g], p->group_labels + p->group_offsets[g + 1]);
        c.voxel_size.push_back(p->voxel_size[g]);
    }
    c.voxel_size_map = p->voxel_size_map, c.max_range = p->max_range, c.min_range = p->min_range;
    c.label_max_range = p->label_max_range, c.local_map_range = p->local_map_range;
    c.basic_points_per_voxel = p->basic_points_per_voxel, c.critical_points_per_voxel = p->critical_points_per_voxel;
    c.basic_parts_labels.assign(p->basic_parts_labels, p->basic_parts_labels + p->n_basic_parts_labels);
    c.min_motion_th = p->min_motion_th, c.initial_threshold = p->initial_threshold, c.sem_th = p->sem_th;
    c.deskew = p->deskew != 0, c.dynamic_vehicle_filter = p->dynamic_vehicle_filter != 0;
    c.dynamic_vehicle_filter_th = p->dynamic_vehicle_filter_th, c.dynamic_vehicle_voxid = p->dynamic_vehicle_voxid;
    c.dynamic_remove_lankmark.assign(p->dynamic_remove_lankmark, p->dynamic_remove_lankmark + p->n_dynamic_remove_lankmark);
    return c;
}

Cloud to_cloud(const double *xyzl, size_t n) {
    Cloud c(n);
    for (size_t i = 0; i < n; ++i) c[i] = Eigen::Vector4d(xyzl[4 * i], xyzl[4 * i + 1], xyzl[4 * i + 2], xyzl[4 * i + 3]);
    return c;
}
size_t from_cloud(const Cloud &c, double *out, size_t cap) {
    if (out)
        for (size_t i = 0; i < c.size() && i < cap; ++i)
            for (int k = 0; k < 4; ++k) out[4 * i + k] = c[i][k];
    return c.size();
}
Sophus::SE3d to_se3(const double p[7]) {  // wire order: tx ty tz qx qy qz qw
    orc::SE3 T;
    T.t = orc::Vec3{p[0], p[1], p[2]};
    T.q = orc::quat_normalized(orc::Quat{p[6], p[3], p[4], p[5]});
    return Sophus::SE3d(T);
}
void from_se3(const Sophus::SE3d &S, double p[7]) {
    const orc::SE3 &T = S.raw();
    p[0] = T.t.x, p[1] = T.t.y, p[2] = T.t.z, p[3] = T.q.x, p[4] = T.q.y, p[5] = T.q.z, p[6] = T.q.w;
}

struct RefPipeline {
    sage_icp::pipeline::sageICP icp;
    Cloud last_source;
    explicit RefPipeline(const sage_icp::pipeline::sageConfig &c) : icp(c) {}
};

}  // namespace

extern "C" {

// ---- free functions ---------------------------------------------------------------------------------------
size_t ref_preprocess(const ref_config_pod *cfg, const double *xyzl, size_t n, double *out, size_t cap) {
    const auto c = to_config(cfg);
    return from_cloud(sage_icp::Preprocess(to_cloud(xyzl, n), c.max_range, c.min_range, c.label_max_range, c.dynamic_vehicle_filter,
                                           c.dynamic_vehicle_filter_th, c.voxel_labels[c.dynamic_vehicle_voxid], c.dynamic_remove_lankmark),
                      out, cap);
}
size_t ref_voxel_downsample(const ref_config_pod *cfg, const double *xyzl, size_t n, double vox_scale, double *out, size_t cap) {
    const auto c = to_config(cfg);
    return from_cloud(sage_icp::VoxelDownsample(to_cloud(xyzl, n), c.voxel_labels, c.voxel_size, vox_scale), out, cap);
}
size_t ref_deskew(const double *frame, const double *ts, size_t n, const double start[7], const double finish[7], double *out) {
    return from_cloud(sage_icp::DeSkewScan(to_cloud(frame, n), std::vector<double>(ts, ts + n), to_se3(start), to_se3(finish)), out, n);
}

// ---- sage_icp::VoxelHashMap -------------------------------------------------------------------------------
void *ref_map_create(double voxel_size, double max_distance, int basic, int critical, const int32_t *labels, int n_labels) {
    return new sage_icp::VoxelHashMap(voxel_size, max_distance, basic, critical, std::vector<int>(labels, labels + n_labels));
}
void ref_map_destroy(void *m) { delete (sage_icp::VoxelHashMap *)m; }
void ref_map_clear(void *m) { ((sage_icp::VoxelHashMap *)m)->Clear(); }
int ref_map_empty(void *m) { return ((sage_icp::VoxelHashMap *)m)->Empty() ? 1 : 0; }
size_t ref_map_num_voxels(void *m) { return ((sage_icp::VoxelHashMap *)m)->map_.size(); }
void ref_map_add_points(void *m, const double *xyzl, size_t n) { ((sage_icp::VoxelHashMap *)m)->AddPoints(to_cloud(xyzl, n)); }
void ref_map_remove_far(void *m, const double origin[3]) {
    ((sage_icp::VoxelHashMap *)m)->RemovePointsFarFromLocation(Eigen::Vector3d(origin[0], origin[1], origin[2]));
}
void ref_map_update(void *m, const double *xyzl, size_t n, const double pose[7]) {
    ((sage_icp::VoxelHashMap *)m)->Update(to_cloud(xyzl, n), to_se3(pose));
}
size_t ref_map_pointcloud(void *m, double *out, size_t cap) { return from_cloud(((sage_icp::VoxelHashMap *)m)->Pointcloud(), out, cap); }
// voxels in the map's iteration order: keys (3 ints), point counts, points (stride x 4 doubles per voxel, zero padded)
size_t ref_map_dump(void *m, int32_t *keys, int32_t *counts, double *pts, int stride, size_t cap_voxels) {
    const auto &map = ((sage_icp::VoxelHashMap *)m)->map_;
    size_t v = 0;
    for (const auto &[voxel, block] : map) {
        if (v >= cap_voxels) break;
        for (int k = 0; k < 3; ++k) keys[3 * v + k] = voxel[k];
        counts[v] = (int32_t)block.points.size();
        std::memset(pts + v * (size_t)stride * 4, 0, sizeof(double) * 4 * (size_t)stride);
        for (size_t j = 0; j < block.points.size() && (int)j < stride; ++j)
            for (int k = 0; k < 4; ++k) pts[(v * (size_t)stride + j) * 4 + k] = block.points[j][k];
        ++v;
    }
    return v;
}
size_t ref_map_get_correspondences(void *m, const double *xyzl, size_t n, double max_dist, double th, double *src_out, double *tgt_out) {
    const auto [src, tgt] = ((sage_icp::VoxelHashMap *)m)->GetCorrespondences(to_cloud(xyzl, n), max_dist, th);
    from_cloud(src, src_out, n);
    from_cloud(tgt, tgt_out, n);
    return src.size();
}
// sage_icp::RegisterFrame (core/Registration.cpp:113-141): 500 iterations at most, |log(est)| < 1e-4 — both fixed in the reference
void ref_register_frame_core(void *m, const double *frame, size_t n, const double guess[7], double max_dist, double kernel, double sem_th,
                             double pose_out[7]) {
    from_se3(sage_icp::RegisterFrame(to_cloud(frame, n), *(sage_icp::VoxelHashMap *)m, to_se3(guess), max_dist, kernel, sem_th), pose_out);
}

// ---- sage_icp::pipeline::sageICP --------------------------------------------------------------------------
void *ref_create(const ref_config_pod *cfg) { return new RefPipeline(to_config(cfg)); }
void ref_destroy(void *h) { delete (RefPipeline *)h; }
void ref_reset(void *h) { ((RefPipeline *)h)->icp.reinitialize(); }
void ref_register_frame(void *h, const double *xyzl, size_t n, const double *ts, double pose_out[7]) {
    auto *p = (RefPipeline *)h;
    const Cloud frame = to_cloud(xyzl, n);
    if (ts)
        p->last_source = std::get<0>(p->icp.RegisterFrame(frame, std::vector<double>(ts, ts + n)));
    else
        p->last_source = std::get<0>(p->icp.RegisterFrame(frame));
    from_se3(p->icp.poses().back(), pose_out);
}
size_t ref_voxelize(void *h, const double *xyzl, size_t n, double *source_out, size_t *n_source, double *ds_out, size_t *n_ds) {
    const auto [source, ds] = ((RefPipeline *)h)->icp.Voxelize(to_cloud(xyzl, n));
    *n_source = from_cloud(source, source_out, n);
    *n_ds = from_cloud(ds, ds_out, n);
    return *n_source;
}
double ref_get_adaptive_threshold(void *h) { return ((RefPipeline *)h)->icp.GetAdaptiveThreshold(); }
int ref_has_moved(void *h) { return ((RefPipeline *)h)->icp.HasMoved() ? 1 : 0; }
void ref_get_prediction_model(void *h, double out[7]) { from_se3(((RefPipeline *)h)->icp.GetPredictionModel(), out); }
size_t ref_last_source(void *h, double *out, size_t cap) { return from_cloud(((RefPipeline *)h)->last_source, out, cap); }
size_t ref_num_poses(void *h) { return ((RefPipeline *)h)->icp.poses().size(); }
void ref_get_pose(void *h, size_t i, double out[7]) { from_se3(((RefPipeline *)h)->icp.poses()[i], out); }
size_t ref_local_map(void *h, double *out, size_t cap) { return from_cloud(((RefPipeline *)h)->icp.LocalMap(), out, cap); }
void ref_transform_to_last_frame(void *h, const double last[7], const double cur[7], const double *xyzl, size_t n, double *out) {
    from_cloud(((RefPipeline *)h)->icp.TransformToLastFrame(to_se3(last), to_se3(cur), to_cloud(xyzl, n)), out, n);
}

}  // extern "C"
