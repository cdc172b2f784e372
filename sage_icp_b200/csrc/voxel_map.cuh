// Device-resident semantic voxel map + registration driver (host class).
// Mirrors sage_icp::VoxelHashMap (core/VoxelHashMap.hpp:34-106) and sage_icp::RegisterFrame
// (core/Registration.cpp:113-141).
#pragma once
#include <vector>

#include "common.cuh"

namespace sage {

struct NcclComm;  // nccl_shim.cu
struct IterParams;  // registration.cu

// Raw device view handed to kernels.
struct MapView {
    TblEntry *tbl;
    uint32_t mask;  // capacity - 1
    unsigned long long *blk_key;
    int32_t *blk_cnt;
    uint32_t *blk_head;  // per-block arrival list head (kNil between updates)
    uint32_t *blk_slot;  // table slot of the block
    double4 *blk_pts;    // [block][stride] x, y, z, label — bit-identical to Eigen::Vector4d ("cold" exact copy)
    float4 *blk_hot;     // [block][stride] 16-byte search record: f32 offsets from the voxel origin + label (see hot_record)
    uint32_t *free_list;
    uint32_t *blk_first;  // faithful-eviction mode only (else null): smallest point index of the batch that created the block,
    uint8_t *blk_new;     // and the "created by this batch" mark
    MapCtrl *ctrl;
    int stride;  // basic + critical
    int basic, critical;
    int n_basic_labels;
    int basic_labels[32];
    double voxel_size;
    uint32_t blk_cap;
};

// Host mirror of the reference's tsl::robin_map<Voxel, VoxelBlock, VoxelHash> as far as it is observable: which bucket
// every voxel sits in.  Kept only in faithful-eviction mode, where RemovePointsFarFromLocation must reproduce the
// erase-while-iterating sweep of core/VoxelHashMap.cpp:176-184 and LocalMap()/dump() the map's iteration order
// (core/VoxelHashMap.cpp:132-142).  Same published rules as frontend.cu's replay, plus backward-shift erase.
uint32_t reference_voxel_hash(unsigned long long packed_key);  // the reference's 20-bit VoxelHash of a packed key

class HostVoxelTable {
public:
    struct Bucket {
        int32_t dist = -1;  // -1 = empty
        uint32_t hash = 0, block = 0;
        unsigned long long key = 0;
    };
    void clear();  // tsl clear(): empties the buckets, keeps the bucket count
    void insert(unsigned long long key, uint32_t hash20, uint32_t block);  // key must be absent
    void erase_at(size_t i);
    size_t bucket_count() const { return b_.size(); }
    size_t size() const { return n_; }
    const Bucket &at(size_t i) const { return b_[i]; }

private:
    void place(std::vector<Bucket> &t, Bucket e, bool track);
    std::vector<Bucket> b_;
    size_t n_ = 0, threshold_ = 0;
    bool grow_next_ = false;
};

class VoxelMapGPU {
public:
    VoxelMapGPU(double voxel_size, double max_distance, int basic, int critical, const int32_t *labels, int n_labels, int device);
    ~VoxelMapGPU();

    void clear();
    bool empty();
    long long num_voxels();
    long long num_points();

    // points already on the device, n x double4; pose == nullptr -> points are already in the map frame
    void add_points_dev(const double4 *pts, size_t n, const Pose *pose);
    void add_points_host(const double *xyzl, size_t n, const Pose *pose);
    void remove_far(double ox, double oy, double oz);
    void update_dev(const double4 *pts, size_t n, const Pose &pose) {
        add_points_dev(pts, n, &pose);
        remove_far(pose.tx, pose.ty, pose.tz);
    }
    long long pointcloud(double *out, size_t cap_points);
    // 0: every far voxel is evicted (default); 1: the reference's erase-while-iterating sweep, bucket for bucket.
    // Only on an empty map.
    void set_eviction_faithful(bool on);
    bool eviction_faithful() const { return faithful_; }
    void load(const int32_t *keys, const int32_t *counts, const double *points, int stride, size_t n_voxels);
    long long dump(int32_t *keys, int32_t *counts, double *points, size_t cap_voxels);

    // Registration on a device-resident frame.  Returns iterations executed.
    int register_frame_dev(const double4 *frame, size_t n, const Pose &guess, double max_dist, double kernel, double sem_th,
                           int max_iters, double est_th, Pose &pose_out);
    int register_frame_host(const double *xyzl, size_t n, const Pose &guess, double max_dist, double kernel, double sem_th,
                            int max_iters, double est_th, Pose &pose_out);
    long long get_correspondences(const double *xyzl, size_t n, double max_dist, double th, double *target_out, uint8_t *matched_out);
    void normal_equations(const double *xyzl, size_t n, double max_dist, double kernel, double sem_th, double JTJ[36], double JTr[6],
                          long long *pairs);
    void nn_stats(const double *xyzl, size_t n, unsigned long long *occupied, unsigned long long *candidates);
    // work the search kernel actually does on these queries: records scanned, table probes, queries re-ranked in f64
    void search_work(const double *xyzl, size_t n, double max_dist, double sem_th, unsigned long long *scanned,
                     unsigned long long *probes, unsigned long long *exact, unsigned long long *heavy = nullptr,
                     unsigned long long *staged = nullptr);

    cudaStream_t stream() const { return stream_; }
    int device() const { return device_; }
    size_t debug_timeline(unsigned long long *out, size_t cap);
    void profile_enable(bool on);
    void profile_read(long long *iterations, double *ms, long long *kernel_launches = nullptr);

    void comm_init(int rank, int world, const uint8_t id[128]);
    void comm_destroy();
    // fused all-reduce over peer memory: export this rank's exchange buffer, then attach every rank's
    void peer_handle(uint8_t out[64]);
    void peer_attach(int rank, int world, const uint8_t *handles);
    void peer_detach();
    // the same exchange between several maps of ONE process (sage_set_devices with n > 1): the buffers are plain device pointers,
    // reachable from every device once peer access is enabled — no IPC handles
    double *peer_local_buffer();
    void peer_attach_local(int rank, int world, double *const *buffers, const int *devices);

    // statistics mirror (refreshed by sync_stats)
    void sync_stats();
    MapCtrl stats() const { return host_stats_; }
    double voxel_size() const { return voxel_size_; }
    int stride() const { return stride_; }

    // scratch upload helper shared with the pipeline: copies n x 4 doubles to a device staging buffer
    double4 *stage_points(const double *xyzl, size_t n);

private:
    MapView view();
    void reserve(size_t extra_points);  // make room for up to `extra_points` new voxels
    void rebuild_table(uint32_t new_cap);
    void launch_iteration(double4 *src, size_t n, double max_dist, double kernel, double sem_th, int mode, double4 *tgt_out,
                          uint8_t *matched_out, int persistent_iters = 0, int iter_index = 0, bool pre_transformed = false);
    void fill_params(IterParams &p, double4 *src, size_t n, double max_dist, double kernel, double sem_th, int mode, double4 *tgt_out,
                     uint8_t *matched_out);
    // tile search (search_tile.cuh, tile_sort.cu): sort + unit list once per registration, then one launch per iteration or one
    // cooperative launch for the whole loop
    void tile_prepare(const double4 *frame, size_t n, const Pose &guess, bool apply_guess, bool with_init = false, int max_iters = 0,
                      double est_th = 0.0);
    int tile_prepare_enqueue(size_t n, bool with_init);  // returns the kernels launched
    void tile_graph_drop();
    void launch_tile(size_t n, double max_dist, double kernel, double sem_th, int iter_index, int persistent_iters);
    void prof_begin();
    void prof_end(int iterations);
    void init_search_config();
    void set_device() const { SAGE_CUDA(cudaSetDevice(device_)); }

    double voxel_size_, max_distance_;
    int basic_, critical_, stride_;
    std::vector<int> basic_labels_;
    int device_;
    cudaStream_t stream_ = nullptr;
    int sm_count_ = 148;

    DevBuf<TblEntry> tbl_;
    uint32_t tbl_cap_ = 0;
    DevBuf<unsigned long long> blk_key_;
    DevBuf<int32_t> blk_cnt_;
    DevBuf<uint32_t> blk_head_, blk_slot_, free_list_;
    DevBuf<double4> blk_pts_;
    DevBuf<float4> blk_hot_;
    uint32_t blk_cap_ = 0;
    DevBuf<MapCtrl> ctrl_;
    PinBuf<MapCtrl> ctrl_pin_;
    MapCtrl host_stats_{};
    size_t hi_bound_ = 0;    // upper bound of ctrl.n_hi
    size_t live_bound_ = 0;  // upper bound of ctrl.n_live

    // update scratch
    DevBuf<double4> upd_pts_;
    DevBuf<uint32_t> upd_slot_, upd_next_;
    // faithful-eviction mode
    bool faithful_ = false;
    HostVoxelTable host_tbl_;
    DevBuf<uint32_t> blk_first_, new_list_, evict_list_;
    DevBuf<uint8_t> blk_new_, far_flags_;
    PinBuf<uint32_t> new_pin_, evict_pin_;
    PinBuf<uint8_t> far_pin_;
    void record_new_voxels(uint32_t m);
    void remove_far_faithful(double ox, double oy, double oz);
    // registration scratch
    DevBuf<double4> stage_, src_, tgt_;
    DevBuf<uint8_t> matched_;
    DevBuf<IcpState> icp_;
    PinBuf<IcpState> icp_pin_;
    DevBuf<double> partials_;
    int nn_grid_ = 0;
    size_t all_warp_max_ = 0;  // scans up to this many queries use the warp-per-query mode
    int persistent_grid_ = 0;    // co-resident blocks of the persistent kernel
    int persistent_grid_wide_ = 0, persistent_grid_widest_ = 0;  // ... of its 128- and 255-register instantiations (small scans)
    size_t persistent_max_ = 0;  // scans up to this many queries run the whole GN loop in one cooperative launch
    int last_iters_ = 0;  // iterations of the previous registration (sizes the first launch batch)
    // tile search
    DevBuf<uint32_t> tile_keys_[2], tile_vals_[2], tile_units_, tile_heads_, tile_nunits_, tile_order_;
    DevBuf<uint8_t> tile_tmp_, tile_flag_;  // sort scratch; unit-head flag per sorted position
    DevBuf<double> tile_unit_part_;      // [unit][17] sums of one unit
    DevBuf<uint32_t> tile_group_cnt_;    // units of a group that have published their sums
    DevBuf<uint32_t> tile_ctl_;          // unit hand-out counter and its per-iteration base, each on its own cache line
    int tile_grid_ = 0;            // co-resident blocks of the tile kernels
    int tile_minb_ = 6;            // which instantiation: 6 (80 registers) or 4 (128 registers) resident blocks per SM aimed at
    size_t tile_min_ = 0;          // scans of at least this many queries take the tile search (0 = never)
    uint32_t tile_stage_cap_ = 0;  // staging area of a block, in 16-byte records
    bool tile_persistent_ = true;  // whole GN loop in one cooperative launch
    bool tile_by_size_ = true;     // hand the units out largest first
    int step_everywhere_ = 1;      // persistent kernels, single rank: every block takes the Gauss-Newton step (0 never, 1 where measured to pay, 2 always)
    size_t tile_fill_ = 1;         // 1: thinly spread query sets go to the per-query kernel (tile_units_too_thin); 0: never
    bool tile_units_too_thin(uint32_t n_units, size_t n) const;
    uint32_t last_units_ = 0;      // units of the last sorted scan
    PinBuf<uint32_t> tile_nunits_pin_;
    // per-call arguments of the sort / unit-list kernels (device copy + pinned source) and the captured graph of those kernels
    DevBuf<TilePrepArgs> prep_args_;
    PinBuf<TilePrepArgs> prep_pin_;
    bool tile_graph_ = true;              // SAGE_TILE_GRAPH=0: plain launches
    cudaGraphExec_t prep_exec_ = nullptr;
    std::vector<const void *> prep_key_, prep_seen_;  // what the graph has baked in / what the previous plain call looked like
    long long prep_kernel_nodes_ = 0;
    size_t tile_tmp_bytes_ = 0, tile_tmp_n_ = (size_t)-1;
    bool coop_ok_ = false;
    int light_probes_ = -1;  // < 0: chosen from the number of queries (launch_iteration); SAGE_LIGHT_PROBES overrides
    bool dbg_on_ = false;
    DevBuf<unsigned long long> dbg_;

    // profiling
    bool profile_ = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events_;
    std::vector<int> prof_iters_;  // Gauss-Newton iterations each timed launch ran (1, or the whole loop of a persistent launch)
    size_t prof_used_ = 0;

    NcclComm *comm_ = nullptr;
    double *peer_buf_[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // [rank] -> mapped exchange buffer
    double *xchg_local_ = nullptr;
    int peer_rank_ = 0, peer_world_ = 0;
    bool peer_inprocess_ = false;  // peer_buf_ are other maps' buffers of this process (nothing to close on detach)
    unsigned long long xchg_tag_ = 0;                       // exchanges completed (same on every rank)
    unsigned long long xchg_timeout_ns_ = 30000000000ull;  // SAGE_XCHG_TIMEOUT_S
};

}  // namespace sage
