/*
 * sage_icp_b200 — C ABI of the B200-native SAGE-ICP registration hot path.
 *
 * Plain C, POD-only: no Eigen / Sophus / torch types cross this boundary.  Every entry point names the reference
 * interface it replaces (paths relative to the reference repo, cpp/sage_icp/...).  The C++ adaptor with the
 * reference's exact public surface lives in include/sage_icp/pipeline/sageICP.hpp and is a thin pimpl over
 * these calls; INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions
 *   - points:  double[n][4] = x, y, z, label — bit-identical to std::vector<Eigen::Vector4d>::data().
 *   - poses:   double[7]    = tx, ty, tz, qx, qy, qz, qw  (Sophus::SE3d: translation + unit quaternion).
 *   - return:  0 / a count on success, negative SAGE_E* on failure; sage_last_error() gives the message.
 *   - All compute runs on the handle's CUDA device.  There is NO CPU fallback: without a usable sm_100 device
 *     sage_create()/sage_map_create() fail with SAGE_ENODEVICE.
 *   - A handle is not thread-safe (the reference class is driven from one rclcpp executor thread,
 *     ros/ros2/OdometryServer.cpp:356).
 */
#ifndef SAGE_ICP_B200_H_
#define SAGE_ICP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SAGE_OK 0
#define SAGE_EINVAL -1    /* bad argument / config (e.g. no voxel groups: reference UB) */
#define SAGE_ENODEVICE -2 /* no CUDA device / wrong architecture */
#define SAGE_ECUDA -3     /* CUDA runtime error (message in sage_last_error) */
#define SAGE_ERANGE -4    /* coordinate outside the packable voxel-key range */
#define SAGE_ENCCL -5     /* NCCL unavailable or failed */
#define SAGE_ECAPACITY -6 /* output buffer too small (needed count is returned through the out-param) */

/* sage_icp::pipeline::sageConfig — pipeline/sageICP.hpp:39-65 (same fields; 2-D voxel_labels flattened). */
typedef struct sage_config_pod {
    int32_t n_groups;             /* voxel_labels.size() == voxel_size.size() */
    const int32_t *group_offsets; /* n_groups+1 offsets into group_labels */
    const int32_t *group_labels;
    const double *voxel_size; /* n_groups */
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
} sage_config_pod;

typedef struct sage_pipeline sage_pipeline; /* sage_icp::pipeline::sageICP */
typedef struct sage_map sage_map;           /* sage_icp::VoxelHashMap      */

const char *sage_last_error(void);
/* Number of visible CUDA devices with compute capability 10.x (0 => every create call fails loudly). */
int sage_device_count(void);

/* ------------------------------------------------------------------------------------------------------------
 * Pipeline level — sage_icp::pipeline::sageICP (pipeline/sageICP.hpp:67-109)
 * ---------------------------------------------------------------------------------------------------------- */

/* sageICP::sageICP(const sageConfig&) — pipeline/sageICP.hpp:73-76.  `device` = CUDA ordinal. */
sage_pipeline *sage_create(const sage_config_pod *config, int device);
void sage_destroy(sage_pipeline *h);
/* sageICP::reinitialize() — pipeline/sageICP.hpp:94-99 */
int sage_reset(sage_pipeline *h);
/* Device selection of the drop-in (the reference is CPU only).  Only on a fresh or reinitialised pipeline.
 *   n == 1: the pipeline runs on GPU ids[0].
 *   n >  1: ONE process, n GPUs of one node (peer access required).  The front end runs on ids[0]; every GPU holds a replica of the
 *           map; RegisterFrame cuts its ICP queries into n contiguous slices, registers them concurrently with the 17
 *           normal-equation sums all-reduced inside the search kernel over NVLink peer memory, checks that every GPU returned the
 *           same pose, and applies the same map update on every replica.  Poses equal the single-GPU ones to rounding (other
 *           summation order).  Changes made through sage_pipeline_map() (loads, eviction mode) reach the first GPU only: set them
 *           before this call.  The ROS node uses several GPUs with this one line after constructing sageICP. */
int sage_set_devices(sage_pipeline *h, const int *ids, int n);
int sage_num_devices(sage_pipeline *h);

/* sageICP::RegisterFrame(frame[, timestamps]) — pipeline/sageICP.cpp:36-52 and :54-95.
 * xyzl: HOST pointer, n x 4 doubles.  timestamps: NULL or n doubles (used only when config.deskew).
 * Outputs: new pose, t_icp / t_all in seconds (same meaning as the reference tuple: t_all excludes the map update).
 * The returned `source` cloud of the reference tuple is fetched with sage_last_source(). */
int sage_register_frame(sage_pipeline *h, const double *xyzl, size_t n, const double *timestamps, double pose_out[7],
                        double *t_icp, double *t_all);
/* utils::PointCloud2ToEigen (ros/ros2/Utils.hpp:161-180) fused with RegisterFrame: `data` is the HOST data buffer of a
 * sensor_msgs/PointCloud2 (width*height = n_points records of point_step bytes; the reference's publishers send packed
 * 17-byte records f32 x,y,z @0/4/8, u8 label @12, u32 rgb @13 — eval/kitti_pub.py:184-207).  The buffer crosses PCIe as
 * is and is widened to f64 on the device, exactly as the reference widens it on the host.  label_datatype follows
 * sensor_msgs/PointField: 2 = UINT8 (messages with 5 fields), 7 = FLOAT32 (the reference's other branch). */
int sage_register_frame_pointcloud2(sage_pipeline *h, const uint8_t *data, size_t n_points, uint32_t point_step,
                                    uint32_t x_offset, uint32_t y_offset, uint32_t z_offset, uint32_t label_offset,
                                    int label_datatype, const double *timestamps, double pose_out[7], double *t_icp,
                                    double *t_all);
/* std::get<0>(RegisterFrame(...)) — the double-downsampled query cloud, sensor frame, reference order. */
int64_t sage_last_source(sage_pipeline *h, double *out, size_t cap_points);
/* frame_downsample of the last frame (what Update() consumed) — diagnostic, pipeline/sageICP.cpp:68,92 */
int64_t sage_last_frame_downsample(sage_pipeline *h, double *out, size_t cap_points);
/* Iterations the last ICP ran and the sigma it used — diagnostics for parity tests. */
int sage_last_iterations(sage_pipeline *h);
double sage_last_sigma(sage_pipeline *h);

/* sageICP::Voxelize — pipeline/sageICP.cpp:97-101.  Returns counts through n_source / n_downsample. */
int sage_voxelize(sage_pipeline *h, const double *xyzl, size_t n, double *source_out, size_t *n_source,
                  double *downsample_out, size_t *n_downsample);
/* sageICP::GetAdaptiveThreshold (stateful, pipeline/sageICP.cpp:103-108), HasMoved (:117-121),
 * GetPredictionModel (:110-115) */
double sage_get_adaptive_threshold(sage_pipeline *h);
int sage_has_moved(sage_pipeline *h);
int sage_get_prediction_model(sage_pipeline *h, double pose_out[7]);
/* sageICP::TransformToLastFrame — pipeline/sageICP.cpp:123-129 (points in/out may alias) */
int sage_transform_to_last_frame(sage_pipeline *h, const double last_pose[7], const double current_pose[7],
                                 const double *xyzl, size_t n, double *out);
/* sageICP::poses() — pipeline/sageICP.hpp:93 */
/* The node's key-frame test on the device (ros/ros2/OdometryServer.cpp:222-241):
 *   grid_out[rows][cols] = utils::EigenToGridMap(points, bounds, {rows, cols})              ros/ros2/Utils.hpp:220-242
 *   with last_pose/current_pose (both or neither): of TransformToLastFrame(last, current, points) pipeline/sageICP.cpp:123-129
 *   with last_occ: *overlap_out = utils::compute_occ_overlap(last_occ, that grid)           ros/ros2/Utils.hpp:244-258
 * bounds = {x0, x1, y0, y1, z0, z1} (key_frame_bounds flattened); grids are row-major int32 0/1 like the reference's
 * vector<vector<int>>; grid_out may be NULL when only the overlap is wanted.  The points cross PCIe once. */
int sage_key_frame_grid(sage_pipeline *h, const double *xyzl, size_t n, const double *last_pose, const double *current_pose,
                        const double bounds[6], int rows, int cols, const int32_t *last_occ, int32_t *grid_out, double *overlap_out);
int64_t sage_num_poses(sage_pipeline *h);
int sage_get_pose(sage_pipeline *h, size_t i, double pose_out[7]);
/* poses()[first ...] in one call: writes min(cap, num_poses - first) poses of 7 doubles each and returns that number
 * (poses_out == NULL or cap == 0: returns how many there are from `first` on).  The ROS node reads poses() after every scan
 * (ros/ros2/OdometryServer.cpp:173): with `first` = the number it already holds, that is one call and one pose per frame
 * instead of one call per pose of the whole drive. */
int64_t sage_get_poses(sage_pipeline *h, size_t first, double *poses_out, size_t cap);
/* sageICP::LocalMap() — pipeline/sageICP.hpp:92 -> VoxelHashMap::Pointcloud (core/VoxelHashMap.cpp:132-142).
 * Same point set as the reference; order is device block order, not robin_map iteration order. */
int64_t sage_local_map(sage_pipeline *h, double *out, size_t cap_points);
/* The pipeline's map (borrowed; do not destroy). */
sage_map *sage_pipeline_map(sage_pipeline *h);

/* ------------------------------------------------------------------------------------------------------------
 * Core level — sage_icp::VoxelHashMap (core/VoxelHashMap.hpp) and the free functions of core/
 * ---------------------------------------------------------------------------------------------------------- */

/* VoxelHashMap::VoxelHashMap — core/VoxelHashMap.hpp:79-88 */
sage_map *sage_map_create(double voxel_size, double max_distance, int basic_points_per_voxel,
                          int critical_points_per_voxel, const int32_t *basic_parts_labels, int n_labels, int device);
void sage_map_destroy(sage_map *m);
int sage_map_clear(sage_map *m);                   /* VoxelHashMap::Clear  — core/VoxelHashMap.hpp:93 */
/* Eviction / iteration-order mode (empty map only).  0 (default): RemovePointsFarFromLocation drops EVERY voxel whose first
 * point is farther than max_distance, and dump/pointcloud come in device block order.  1: bucket-for-bucket emulation of the
 * reference's erase-while-iterating sweep over its tsl::robin_map (core/VoxelHashMap.cpp:176-184 skips the element that
 * backward-shift deletion moves into the erased bucket), and dump/pointcloud/LocalMap() in that map's iteration order
 * (core/VoxelHashMap.cpp:132-142).  Mode 1 keeps a host mirror of the bucket array and adds host round trips per update. */
int sage_map_set_eviction(sage_map *m, int faithful);
int sage_map_empty(sage_map *m);                   /* VoxelHashMap::Empty  — core/VoxelHashMap.hpp:94 */
int64_t sage_map_num_voxels(sage_map *m);
int64_t sage_map_num_points(sage_map *m);
/* VoxelHashMap::AddPoints — core/VoxelHashMap.cpp:162-174 (exact sequential AddPoint semantics per voxel) */
int sage_map_add_points(sage_map *m, const double *xyzl, size_t n);
/* VoxelHashMap::RemovePointsFarFromLocation — core/VoxelHashMap.cpp:176-184 ("clean" eviction: every voxel whose
 * first point is farther than max_distance goes; see DESIGN.md on the reference's erase-while-iterating skip) */
int sage_map_remove_far(sage_map *m, const double origin[3]);
/* VoxelHashMap::Update(points, pose) — core/VoxelHashMap.cpp:149-160 */
int sage_map_update(sage_map *m, const double *xyzl, size_t n, const double pose[7]);
/* VoxelHashMap::Pointcloud — core/VoxelHashMap.cpp:132-142 */
int64_t sage_map_pointcloud(sage_map *m, double *out, size_t cap_points);
/* Bulk load of a pre-built map (voxel keys V x 3, per-voxel counts, points V x stride x 4 in stored order). */
int sage_map_load(sage_map *m, const int32_t *keys, const int32_t *counts, const double *points, int stride,
                  size_t n_voxels);
/* Dump: keys V x 3, counts V, points V x stride x 4 (stride = basic+critical).  Returns V (call with NULLs to size). */
int64_t sage_map_dump(sage_map *m, int32_t *keys, int32_t *counts, double *points, size_t cap_voxels);

/* VoxelHashMap::GetCorrespondences — core/VoxelHashMap.cpp:48-130.
 * target_out: n x 4 (valid where matched); matched_out: n flags; returns the number of pairs.  Pairs are reported
 * per query index (the reference's TBB order is unspecified). */
int64_t sage_map_get_correspondences(sage_map *m, const double *xyzl, size_t n, double max_correspondance_distance,
                                     double th, double *target_out, uint8_t *matched_out);
/* Exact neighbourhood statistics used for the algorithmic-bytes roofline figure. */
int sage_map_nn_stats(sage_map *m, const double *xyzl, size_t n, uint64_t *occupied_voxels, uint64_t *candidates);
/* Work the search kernels actually perform for these queries (one GetCorrespondences pass): 16-byte records ranked,
 * hash-table probes, queries whose f32 ranking was ambiguous and were re-ranked with the exact f64 scan, queries the
 * thread-per-query phase handed to a whole warp, and — tile search only, else 0 — records pulled into shared memory by TMA bulk
 * copies (deferred_queries and staged_records may be NULL). */
int sage_map_search_work(sage_map *m, const double *xyzl, size_t n, double max_correspondance_distance, double th,
                         uint64_t *records_scanned, uint64_t *table_probes, uint64_t *exact_queries,
                         uint64_t *deferred_queries, uint64_t *staged_records);

/* sage_icp::RegisterFrame(frame, voxel_map, initial_guess, max_correspondence_distance, kernel, sem_th)
 * — core/Registration.cpp:113-141.  max_iterations / estimation_threshold default to the reference's constants
 * (500, 1e-4; core/Registration.cpp:96-97) when passed as <= 0 / < 0.  HOST buffers. */
int sage_core_register_frame(sage_map *m, const double *frame_xyzl, size_t n, const double initial_guess[7],
                             double max_correspondence_distance, double kernel, double sem_th, int max_iterations,
                             double estimation_threshold, double pose_out[7], int *iterations_out);
/* Same, DEVICE-resident frame (n x 4 doubles in HBM of the map's device); nothing crosses PCIe but the pose. */
int sage_core_register_frame_device(sage_map *m, const void *frame_xyzl_dev, size_t n, const double initial_guess[7],
                                    double max_correspondence_distance, double kernel, double sem_th,
                                    int max_iterations, double estimation_threshold, double pose_out[7],
                                    int *iterations_out);
/* One AlignClouds evaluation (core/Registration.cpp:59-94) on the current correspondences of `frame` against the
 * map: JTJ (36, row-major), JTr (6), and the number of pairs — for parity tests of the reduction. */
int sage_core_normal_equations(sage_map *m, const double *frame_xyzl, size_t n, double max_correspondence_distance,
                               double kernel, double sem_th, double JTJ_out[36], double JTr_out[6],
                               int64_t *n_pairs_out);

/* sage_icp::Preprocess — core/Preprocessing.cpp:86-189: the range branch (:173-187), or the dynamic-vehicle branch (:95-172)
 * when the pipeline's config has dynamic_vehicle_filter set.  Returns the kept count. */
int64_t sage_preprocess(sage_pipeline *h, const double *xyzl, size_t n, double *out, size_t cap_points);
/* sage_icp::VoxelDownsample — core/Preprocessing.cpp:44-84 (reference output order). */
int64_t sage_voxel_downsample(sage_pipeline *h, const double *xyzl, size_t n, double vox_scale, double *out,
                              size_t cap_points);

/* ------------------------------------------------------------------------------------------------------------
 * Measurement and multi-GPU plumbing (no reference counterpart: the reference is single-process CPU)
 * ---------------------------------------------------------------------------------------------------------- */

/* The CUDA stream (cudaStream_t) the map's kernels are launched on, for event timing by the caller. */
void *sage_map_stream(sage_map *m);
/* Per-kernel CUDA-event timing of the correspondence+normal-equation kernel.  enable=1 records an event pair around
 * every launch; sage_map_profile_read() synchronises, returns the Gauss-Newton iterations those launches ran (a launch runs one
 * iteration, or — the cooperative kernels — the whole loop of a registration) and their total milliseconds since the last
 * read; sage_map_profile_read_launches() also returns the number of launches. */
int sage_map_profile_enable(sage_map *m, int enable);
int sage_map_profile_read(sage_map *m, int64_t *iterations, double *total_ms);
int sage_map_profile_read_launches(sage_map *m, int64_t *iterations, int64_t *launches, double *total_ms);
/* Total kernel launches issued by this library in the calling process since load (bench.py's gpu_launches). */
int64_t sage_launch_count(void);

/* Host utility behind sage_voxel_downsample: iteration order of the reference's unreserved tsl::robin_map<Voxel, ...> (v1.0.1)
 * after `n` DISTINCT keys with the given 20-bit VoxelHash values (core/Preprocessing.cpp:35-40) were inserted in this
 * order — core/Preprocessing.cpp:76-82 emits the down-sampled cloud in that order.  order_out[j] = input position of the
 * j-th element.  Pure host code (no device needed); exported so that it can be tested on its own. */
int sage_robin_iteration_order(const uint32_t *hash20, size_t n, uint32_t *order_out);

/* Host mirror table behind sage_map_set_eviction(m, 1), exported so it can be tested without a device: replays n_ops records
 * {op, x, y, z, r2} on an empty table — op 0 inserts voxel (x,y,z) (must be absent), op 1 runs the reference's
 * erase-while-iterating sweep (core/VoxelHashMap.cpp:176-184) erasing every VISITED voxel whose integer squared distance from
 * (x,y,z) exceeds r2, op 2 is clear().  Writes the surviving voxels in iteration (bucket) order to keys_out (3 ints each, at
 * most cap voxels) and the bucket count to *bucket_count (may be NULL).  Returns the number of survivors, or < 0. */
int64_t sage_robin_table_replay(const int32_t *ops, size_t n_ops, int32_t *keys_out, size_t cap, uint64_t *bucket_count);

/* Query shard owned by `rank` out of `world` for n queries: contiguous [begin, end).  Pure host arithmetic. */
int sage_shard_range(size_t n, int rank, int world, size_t *begin, size_t *end);
/* NCCL plumbing for the sharded ICP: rank 0 calls sage_nccl_unique_id and ships the 128 bytes to its
 * peers by any means; every rank then calls sage_map_comm_init.  After that sage_core_register_frame* treats its
 * frame as this rank's shard and all-reduces the normal-equation sums (17 doubles) every iteration. */
int sage_nccl_unique_id(uint8_t id_out[128]);
int sage_map_comm_init(sage_map *m, int rank, int world, const uint8_t id[128]);
/* Fused alternative for the GPUs of ONE NVLink/NVSwitch box (one process per GPU): the all-reduce happens inside the search
 * kernel — every rank's last block stores its 17 sums straight into every peer's exchange buffer (CUDA IPC peer mappings)
 * and adds the slots in rank order, so an iteration stays a single launch.  Every rank calls sage_map_comm_peer_handle,
 * the 64-byte handles are gathered by any means (rank order), then every rank calls sage_map_comm_peer_attach with all of
 * them.  Takes precedence over the NCCL path when both are set up.  A rank that does not arrive within 2 s ends the
 * registration with SAGE_ECUDA. */
int sage_map_comm_peer_handle(sage_map *m, uint8_t handle_out[64]);
int sage_map_comm_peer_attach(sage_map *m, int rank, int world, const uint8_t *handles /* world x 64 bytes */);
int sage_map_comm_destroy(sage_map *m);

#ifdef __cplusplus
}
#endif
#endif /* SAGE_ICP_B200_H_ */
