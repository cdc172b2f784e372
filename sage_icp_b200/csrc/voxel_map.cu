// Device-resident semantic voxel map: open-addressed hash (16-byte entries: key, block id, count) over a pool of
// voxel blocks, each block = `stride` contiguous 32-byte point records (x, y, z, label as f64 — the very bytes of
// the reference's Eigen::Vector4d).  Replaces sage_icp::VoxelHashMap's storage and update path:
//   AddPoints                      core/VoxelHashMap.cpp:162-174  -> insert_keys / link / replay kernels
//   VoxelBlock::AddPoint           core/VoxelHashMap.hpp:45-70    -> add_point_rule(), replayed per voxel in arrival order
//   RemovePointsFarFromLocation    core/VoxelHashMap.cpp:176-184  -> evict kernel + table rebuild
//   Update(points, pose)           core/VoxelHashMap.cpp:149-160  -> transform fused into insert_keys
//   Pointcloud                     core/VoxelHashMap.cpp:132-142  -> pointcloud()
// The reference's sequential insert is order dependent only *within* a voxel, so voxels are replayed in parallel,
// each by one thread walking its arrivals in input order (exact).
#include <algorithm>
#include <array>
#include <cstring>

#include "nccl_shim.cuh"
#include <cstdlib>
#include "voxel_map.cuh"

namespace sage {

std::atomic<long long> g_launches{0};

constexpr int kThreads = 256;
static inline unsigned blocks_for(size_t n, int threads = kThreads) { return (unsigned)((n + threads - 1) / threads); }

// ---------------------------------------------------------------------------------------------
// kernels

__global__ void tbl_clear_kernel(TblEntry *tbl, uint32_t cap, const MapCtrl *ctrl, int only_if_evicted) {
    if (only_if_evicted && ctrl->evicted == 0) return;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cap) tbl[i] = TblEntry{kEmptyKey, kNil, 0};
}

__device__ __forceinline__ uint32_t tbl_insert_slot(TblEntry *tbl, uint32_t mask, unsigned long long key, bool &won) {
    uint32_t i = (uint32_t)mix64(key) & mask;
    won = false;
    while (true) {
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(&tbl[i].key);
        if (cur == key) return i;
        if (cur == kEmptyKey) {
            const unsigned long long old = atomicCAS(&tbl[i].key, kEmptyKey, key);
            if (old == kEmptyKey) {
                won = true;
                return i;
            }
            if (old == key) return i;
        }
        i = (i + 1) & mask;
    }
}

// re-insert every live block (after eviction or growth)
__global__ void tbl_reinsert_kernel(MapView m, int only_if_evicted) {
    if (only_if_evicted && m.ctrl->evicted == 0) return;
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= m.ctrl->n_hi) return;
    const unsigned long long key = m.blk_key[b];
    if (key == kEmptyKey) return;
    bool won;
    const uint32_t s = tbl_insert_slot(m.tbl, m.mask, key, won);
    m.tbl[s].block = b;
    m.tbl[s].count = (uint32_t)m.blk_cnt[b];
    m.blk_slot[b] = s;
}

__global__ void ctrl_finish_kernel(MapCtrl *ctrl) {
    ctrl->evicted = 0;
    if (ctrl->n_free < 0) ctrl->n_free = 0;
}

// phase 1 of AddPoints: (optionally) transform, compute the voxel key, find-or-create the voxel.
__global__ void map_insert_keys_kernel(MapView m, const double4 *in, double4 *pts, uint32_t *slot_out, uint32_t n, int has_pose,
                                       Pose pose) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double4 p = in[i];
    if (has_pose) {  // pose * point, core/VoxelHashMap.cpp:151-157
        double x, y, z;
        pose_act(pose, p.x, p.y, p.z, x, y, z);
        p.x = x, p.y = y, p.z = z;
    }
    pts[i] = p;
    // Voxel((point / voxel_size_).cast<int>()), core/VoxelHashMap.cpp:165
    const int kx = trunc_div(p.x, m.voxel_size), ky = trunc_div(p.y, m.voxel_size), kz = trunc_div(p.z, m.voxel_size);
    if (!key_in_range(kx, ky, kz)) {
        slot_out[i] = kNil;
        atomicAdd(&m.ctrl->range_err, 1u);
        return;
    }
    const unsigned long long key = pack_key(kx, ky, kz);
    bool won;
    const uint32_t s = tbl_insert_slot(m.tbl, m.mask, key, won);
    if (won) {
        const int f = atomicSub(&m.ctrl->n_free, 1);
        uint32_t b;
        if (f > 0)
            b = m.free_list[f - 1];
        else
            b = atomicAdd(&m.ctrl->n_hi, 1u);
        if (b >= m.blk_cap) {
            atomicExch(&m.ctrl->overflow, 1u);
            b = m.blk_cap - 1;
        }
        m.blk_key[b] = key;
        m.blk_cnt[b] = 0;
        m.blk_head[b] = kNil;
        m.blk_slot[b] = s;
        if (m.blk_new) m.blk_new[b] = 1, m.blk_first[b] = kNil;
        m.tbl[s].count = 0;
        m.tbl[s].block = b;
        atomicAdd(&m.ctrl->n_live, 1u);
    }
    slot_out[i] = s;
}

// phase 2: push every point on its voxel's arrival list
__global__ void map_link_kernel(MapView m, const uint32_t *slot, uint32_t *next, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && m.ctrl->n_free < 0) m.ctrl->n_free = 0;
    if (i >= n) return;
    const uint32_t s = slot[i];
    if (s == kNil) {
        next[i] = kNil - 1;  // not on any list
        return;
    }
    const uint32_t b = m.tbl[s].block;
    next[i] = atomicExch(&m.blk_head[b], i);
}

// VoxelBlock::AddPoint — core/VoxelHashMap.hpp:45-70
__device__ __forceinline__ void add_point_rule(const MapView &m, double4 *vox, float4 *hot, int kx, int ky, int kz, int &cnt,
                                               const double4 &p) {
    if (cnt < m.basic || cnt == 0) {  // cnt == 0: a new voxel is created holding its first point whatever the label
        hot[cnt] = hot_record(p, kx, ky, kz, m.voxel_size);
        vox[cnt++] = p;
        return;
    }
    const int label = __double2int_rz(p.w);
    if (label == 0) return;
    bool is_basic = false;
    for (int k = 0; k < m.n_basic_labels; ++k) is_basic |= (m.basic_labels[k] == label);
    if (!is_basic && cnt < m.basic + m.critical) {
        hot[cnt] = hot_record(p, kx, ky, kz, m.voxel_size);
        vox[cnt++] = p;
        return;
    }
    for (int k = 0; k < cnt; ++k)
        if (__double2int_rz(vox[k].w) == 0) {
            hot[k] = hot_record(p, kx, ky, kz, m.voxel_size);
            vox[k] = p;
            return;
        }
}

// Bottom-up merge sort of a voxel's arrival list by point index (the list is built by atomicExch, i.e. in arbitrary order; the
// reference inserts in input order).  O(k log k) pointer steps, no extra memory: a degenerate batch that puts 10^5 points into
// one voxel costs a second, not the hours a quadratic selection would.
__device__ uint32_t sort_arrivals(uint32_t head, uint32_t *next) {
    for (uint32_t insize = 1;; insize *= 2) {
        uint32_t p = head, tail = kNil, nmerges = 0;
        head = kNil;
        while (p != kNil) {
            ++nmerges;
            uint32_t q = p, psize = 0, qsize = insize;
            for (uint32_t i = 0; i < insize; ++i) {
                ++psize;
                q = next[q];
                if (q == kNil) break;
            }
            while (psize > 0 || (qsize > 0 && q != kNil)) {
                uint32_t e;
                if (psize == 0) {
                    e = q, q = next[q], --qsize;
                } else if (qsize == 0 || q == kNil || p <= q) {
                    e = p, p = next[p], --psize;
                } else {
                    e = q, q = next[q], --qsize;
                }
                if (tail != kNil)
                    next[tail] = e;
                else
                    head = e;
                tail = e;
            }
            p = q;
        }
        if (tail != kNil) next[tail] = kNil;
        if (nmerges <= 1) return head;
    }
}

// phase 3: the thread whose point is the head of its voxel's arrival list (the last one pushed) sorts the list into input order
// and replays it.  Ownership is read from blk_head, which only the owner changes (to nil, when it is done): the `next` links are
// rewritten by the sort and must not be used for that.
__global__ void map_replay_kernel(MapView m, const double4 *pts, const uint32_t *slot, uint32_t *next, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = slot[i];
    if (s == kNil) return;  // key outside the packable range: dropped in phase 1
    const uint32_t b = m.tbl[s].block;
    if (m.blk_head[b] != i) return;
    double4 *vox = m.blk_pts + (size_t)b * m.stride;
    float4 *hot = m.blk_hot + (size_t)b * m.stride;
    int kx, ky, kz;
    unpack_key(m.blk_key[b], kx, ky, kz);
    int cnt = m.blk_cnt[b];
    for (uint32_t j = sort_arrivals(m.blk_head[b], next); j != kNil; j = next[j]) add_point_rule(m, vox, hot, kx, ky, kz, cnt, pts[j]);
    m.blk_cnt[b] = cnt;
    m.tbl[s].count = (uint32_t)cnt;
    m.blk_head[b] = kNil;
}

// ---- faithful-eviction mode: the host mirrors the reference's robin table, the device reports what it needs -------------
// order in which this batch created voxels = order of their first point in the batch (core/VoxelHashMap.cpp:163-173)
__global__ void map_mark_first_kernel(MapView m, const uint32_t *slot, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || slot[i] == kNil) return;
    const uint32_t b = m.tbl[slot[i]].block;
    if (m.blk_new[b]) atomicMin(m.blk_first + b, i);
}
__global__ void map_collect_new_kernel(MapView m, const uint32_t *slot, uint32_t n, uint32_t *out, uint32_t *count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || slot[i] == kNil) return;
    const uint32_t b = m.tbl[slot[i]].block;
    if (!m.blk_new[b] || m.blk_first[b] != i) return;
    const uint32_t pos = atomicAdd(count, 1u);
    const unsigned long long key = m.blk_key[b];
    out[4 * pos] = i, out[4 * pos + 1] = b, out[4 * pos + 2] = (uint32_t)key, out[4 * pos + 3] = (uint32_t)(key >> 32);
}
__global__ void map_clear_new_kernel(MapView m, const uint32_t *list, const uint32_t *count) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < *count) m.blk_new[list[4 * i + 1]] = 0;
}
__global__ void map_far_flags_kernel(MapView m, double ox, double oy, double oz, double max_d2, uint8_t *flags, uint32_t n_blocks) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_blocks) return;
    uint8_t far = 0;
    if (b < m.ctrl->n_hi && m.blk_key[b] != kEmptyKey) {
        const double4 f = m.blk_pts[(size_t)b * m.stride];  // voxel_block.points.front()
        const double dx = __dsub_rn(f.x, ox), dy = __dsub_rn(f.y, oy), dz = __dsub_rn(f.z, oz);
        far = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)) > max_d2 ? 1 : 0;
    }
    flags[b] = far;
}
__global__ void map_evict_list_kernel(MapView m, const uint32_t *list, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t b = list[i];
    m.blk_key[b] = kEmptyKey;
    m.blk_cnt[b] = 0;
    const int idx = atomicAdd(&m.ctrl->n_free, 1);
    m.free_list[idx] = b;
    atomicSub(&m.ctrl->n_live, 1u);
    atomicAdd(&m.ctrl->evicted, 1u);
}

// RemovePointsFarFromLocation — core/VoxelHashMap.cpp:176-184, "clean" semantics (every far voxel goes)
__global__ void map_evict_kernel(MapView m, double ox, double oy, double oz, double max_d2) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= m.ctrl->n_hi) return;
    if (m.blk_key[b] == kEmptyKey) return;
    const double4 f = m.blk_pts[(size_t)b * m.stride];  // voxel_block.points.front()
    const double dx = __dsub_rn(f.x, ox), dy = __dsub_rn(f.y, oy), dz = __dsub_rn(f.z, oz);
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    if (d2 > max_d2) {
        m.blk_key[b] = kEmptyKey;
        m.blk_cnt[b] = 0;
        const int idx = atomicAdd(&m.ctrl->n_free, 1);
        m.free_list[idx] = b;
        atomicSub(&m.ctrl->n_live, 1u);
        atomicAdd(&m.ctrl->evicted, 1u);
    }
}

// search records of a bulk-loaded map (load()): one thread per slot
__global__ void map_build_hot_kernel(MapView m, uint32_t n_blocks) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n_blocks * m.stride) return;
    const uint32_t b = (uint32_t)(i / m.stride);
    const int j = (int)(i % m.stride);
    if (m.blk_key[b] == kEmptyKey || j >= m.blk_cnt[b]) return;
    int kx, ky, kz;
    unpack_key(m.blk_key[b], kx, ky, kz);
    m.blk_hot[i] = hot_record(m.blk_pts[i], kx, ky, kz, m.voxel_size);
}

__global__ void map_count_points_kernel(MapView m) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long c = 0;
    if (b < m.ctrl->n_hi && m.blk_key[b] != kEmptyKey) c = (unsigned long long)m.blk_cnt[b];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(&m.ctrl->n_points, c);
}

__global__ void ctrl_set_kernel(MapCtrl *ctrl, uint32_t n_hi, uint32_t n_live) {
    ctrl->n_hi = n_hi, ctrl->n_free = 0, ctrl->n_live = n_live, ctrl->evicted = 0, ctrl->overflow = 0, ctrl->range_err = 0;
    ctrl->n_points = 0;
}

// ---------------------------------------------------------------------------------------------
// host class

static std::vector<int32_t> checked_labels(const int32_t *labels, int n) {
    if (n < 0 || (n > 0 && !labels)) throw ArgError("basic_parts_labels is NULL or its length negative");
    return n ? std::vector<int32_t>(labels, labels + n) : std::vector<int32_t>();
}

void preload_voxel_map_kernels() {
    const void *ks[] = {(const void *)tbl_clear_kernel, (const void *)tbl_reinsert_kernel, (const void *)ctrl_finish_kernel, (const void *)map_insert_keys_kernel, (const void *)map_link_kernel, (const void *)map_replay_kernel, (const void *)map_mark_first_kernel, (const void *)map_collect_new_kernel, (const void *)map_clear_new_kernel, (const void *)map_far_flags_kernel, (const void *)map_evict_list_kernel, (const void *)map_evict_kernel, (const void *)map_build_hot_kernel, (const void *)map_count_points_kernel, (const void *)ctrl_set_kernel};
    for (const void *k : ks) preload_kernel(k);
}

VoxelMapGPU::VoxelMapGPU(double voxel_size, double max_distance, int basic, int critical, const int32_t *labels, int n_labels,
                         int device)
    : voxel_size_(voxel_size), max_distance_(max_distance), basic_(basic), critical_(critical), stride_(basic + critical),
      basic_labels_(checked_labels(labels, n_labels)), device_(device) {
    if (!(voxel_size > 0) || basic < 0 || critical < 0 || basic + critical < 1) throw ArgError("bad voxel map parameters");
    if (n_labels > 32) throw ArgError("at most 32 basic_parts_labels supported");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
        throw CudaError("no usable CUDA device " + std::to_string(device) + " (sage_icp_b200 has no CPU fallback)");
    cudaDeviceProp prop;
    SAGE_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) throw CudaError(std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) + ", need sm_100 (B200)");
    sm_count_ = prop.multiProcessorCount;
    set_device();
    SAGE_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    preload_voxel_map_kernels(), preload_registration_kernels(), preload_tile_sort_kernels();  // no first-use stalls mid-drive
    ctrl_.ensure(1);
    ctrl_pin_.ensure(1);
    icp_.ensure(1);
    icp_pin_.ensure(1);
    clear();
}

VoxelMapGPU::~VoxelMapGPU() {
    cudaSetDevice(device_);
    if (stream_) cudaStreamSynchronize(stream_);
    comm_destroy();
    peer_detach();
    tile_graph_drop();
    if (xchg_local_) cudaFree(xchg_local_);
    for (auto &e : prof_events_) {
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    if (stream_) cudaStreamDestroy(stream_);
}

MapView VoxelMapGPU::view() {
    MapView v;
    v.tbl = tbl_.p, v.mask = tbl_cap_ - 1;
    v.blk_key = blk_key_.p, v.blk_cnt = blk_cnt_.p, v.blk_head = blk_head_.p, v.blk_slot = blk_slot_.p, v.blk_pts = blk_pts_.p, v.blk_hot = blk_hot_.p;
    v.blk_first = faithful_ ? blk_first_.p : nullptr, v.blk_new = faithful_ ? blk_new_.p : nullptr;
    v.free_list = free_list_.p, v.ctrl = ctrl_.p;
    v.stride = stride_, v.basic = basic_, v.critical = critical_;
    v.n_basic_labels = (int)basic_labels_.size();
    for (int i = 0; i < 32; ++i) v.basic_labels[i] = i < (int)basic_labels_.size() ? basic_labels_[i] : 0;
    v.voxel_size = voxel_size_;
    v.blk_cap = blk_cap_;
    return v;
}

void VoxelMapGPU::clear() {
    set_device();
    if (tbl_cap_ == 0) {
        tbl_cap_ = 1u << 16;
        tbl_.ensure(tbl_cap_);
    }
    SAGE_LAUNCH(tbl_clear_kernel, blocks_for(tbl_cap_), kThreads, 0, stream_, tbl_.p, tbl_cap_, ctrl_.p, 0);
    SAGE_LAUNCH(ctrl_set_kernel, 1, 1, 0, stream_, ctrl_.p, 0u, 0u);
    hi_bound_ = live_bound_ = 0;
    host_stats_ = MapCtrl{};
    host_tbl_.clear();
}

void VoxelMapGPU::sync_stats() {
    set_device();
    SAGE_CUDA(cudaMemcpyAsync(ctrl_pin_.p, ctrl_.p, sizeof(MapCtrl), cudaMemcpyDeviceToHost, stream_));
    SAGE_CUDA(cudaStreamSynchronize(stream_));
    host_stats_ = *ctrl_pin_.p;
    hi_bound_ = host_stats_.n_hi;
    live_bound_ = host_stats_.n_live;
    if (host_stats_.overflow) throw CudaError("voxel map pool overflow (internal capacity bound violated)");
}

bool VoxelMapGPU::empty() {
    if (live_bound_ == 0) return true;
    sync_stats();
    return host_stats_.n_live == 0;
}

long long VoxelMapGPU::num_voxels() {
    sync_stats();
    return host_stats_.n_live;
}

long long VoxelMapGPU::num_points() {
    set_device();
    sync_stats();
    if (host_stats_.n_hi == 0) return 0;
    SAGE_CUDA(cudaMemsetAsync(&ctrl_.p->n_points, 0, sizeof(unsigned long long), stream_));
    SAGE_LAUNCH(map_count_points_kernel, blocks_for(host_stats_.n_hi), kThreads, 0, stream_, view());
    sync_stats();
    return (long long)host_stats_.n_points;
}

void VoxelMapGPU::rebuild_table(uint32_t new_cap) {
    if (new_cap != tbl_cap_) {
        tbl_.ensure(new_cap);
        tbl_cap_ = new_cap;
    }
    SAGE_LAUNCH(tbl_clear_kernel, blocks_for(tbl_cap_), kThreads, 0, stream_, tbl_.p, tbl_cap_, ctrl_.p, 0);
    if (hi_bound_ > 0) SAGE_LAUNCH(tbl_reinsert_kernel, blocks_for(hi_bound_), kThreads, 0, stream_, view(), 0);
}

void VoxelMapGPU::reserve(size_t extra) {
    set_device();
    const bool need_blocks = hi_bound_ + extra > blk_cap_;
    const bool need_tbl = 2 * (live_bound_ + extra) > tbl_cap_;
    if (need_blocks || need_tbl) sync_stats();  // tighten the bounds before paying for growth
    static const bool trace = getenv("SAGE_TRACE_GROWTH") != nullptr;
    if (hi_bound_ + extra > blk_cap_) {
        if (trace) fprintf(stderr, "[sage] block pool grows: hi_bound %zu + %zu > cap %u (n_hi %u, live %u)\n", hi_bound_, extra, blk_cap_, host_stats_.n_hi, host_stats_.n_live);
        size_t want = std::max<size_t>(hi_bound_ + extra, (size_t)blk_cap_ * 2);
        // first allocation: room for 64 Ki voxels (126 MB with 40 slots per voxel) so that a streaming map does not pay a
        // reallocate-and-copy step (cudaMalloc + cudaFree: ~10 ms, device-synchronising) every time it doubles early in a drive
        want = std::max<size_t>(want, 65536);
        if (want > 0x7fffffffull / (size_t)stride_) throw ArgError("voxel map too large");
        blk_key_.ensure(want, stream_, true);
        blk_cnt_.ensure(want, stream_, true);
        blk_head_.ensure(want, stream_, true);
        blk_slot_.ensure(want, stream_, true);
        free_list_.ensure(want, stream_, true);
        blk_pts_.ensure(want * (size_t)stride_, stream_, true);
        blk_hot_.ensure(want * (size_t)stride_, stream_, true);
        if (faithful_) {
            const size_t old = blk_new_.cap;
            blk_first_.ensure(want, stream_, true);
            blk_new_.ensure(want, stream_, true);
            SAGE_CUDA(cudaMemsetAsync(blk_new_.p + old, 0, blk_new_.cap - old, stream_));
        }
        blk_cap_ = (uint32_t)std::min<size_t>({blk_key_.cap, blk_cnt_.cap, blk_head_.cap, blk_slot_.cap, free_list_.cap,
                                                blk_pts_.cap / (size_t)stride_, blk_hot_.cap / (size_t)stride_});
    }
    if (2 * (live_bound_ + extra) > tbl_cap_) {
        if (trace) fprintf(stderr, "[sage] table grows: 2 * (live_bound %zu + %zu) > cap %u\n", live_bound_, extra, tbl_cap_);
        uint32_t cap = tbl_cap_;
        while ((size_t)cap < 2 * (live_bound_ + extra)) cap *= 2;
        rebuild_table(cap);
    }
}

void VoxelMapGPU::add_points_dev(const double4 *pts, size_t n, const Pose *pose) {
    set_device();
    const size_t kChunk = 1u << 22;  // bounds scratch; chunks are applied in order so semantics are unchanged
    for (size_t off = 0; off < n; off += kChunk) {
        const uint32_t m = (uint32_t)std::min(kChunk, n - off);
        reserve(m);
        upd_pts_.ensure(m);
        upd_slot_.ensure(m);
        upd_next_.ensure(m);
        MapView v = view();
        SAGE_LAUNCH(map_insert_keys_kernel, blocks_for(m), kThreads, 0, stream_, v, pts + off, upd_pts_.p, upd_slot_.p, m,
                    pose ? 1 : 0, pose ? *pose : pose_identity());
        SAGE_LAUNCH(map_link_kernel, blocks_for(m), kThreads, 0, stream_, v, upd_slot_.p, upd_next_.p, m);
        SAGE_LAUNCH(map_replay_kernel, blocks_for(m), kThreads, 0, stream_, v, upd_pts_.p, upd_slot_.p, upd_next_.p, m);
        hi_bound_ += m;
        live_bound_ += m;
        if (faithful_) record_new_voxels(m);
    }
}

void VoxelMapGPU::add_points_host(const double *xyzl, size_t n, const Pose *pose) {
    const size_t kChunk = 1u << 22;
    for (size_t off = 0; off < n; off += kChunk) {
        const size_t m = std::min(kChunk, n - off);
        double4 *d = stage_points(xyzl + 4 * off, m);
        add_points_dev(d, m, pose);
        SAGE_CUDA(cudaStreamSynchronize(stream_));  // staging buffer is reused by the next chunk
    }
}

// ---------------------------------------------------------------------------------------------
// HostVoxelTable — published rules of tsl::robin_map v1.0.1 (power-of-two growth, robin-hood insertion, backward-shift erase)

void HostVoxelTable::clear() {
    for (auto &e : b_) e = Bucket{};
    n_ = 0, grow_next_ = false;
}

// rehash placement: plain robin-hood walk from the ideal bucket, no growth bookkeeping
void HostVoxelTable::place(std::vector<Bucket> &t, Bucket e, bool) {
    const size_t mask = t.size() - 1;
    size_t i = e.hash & mask;
    for (int d = 0;; ++d, i = (i + 1) & mask) {
        if (d <= t[i].dist) continue;
        e.dist = d;
        if (t[i].dist < 0) {
            t[i] = e;
            return;
        }
        std::swap(e, t[i]);
        d = e.dist;
    }
}

void HostVoxelTable::insert(unsigned long long key, uint32_t hash20, uint32_t block) {
    constexpr int kDistLimit = 8192;  // DIST_FROM_IDEAL_BUCKET_LIMIT
    size_t i = 0;
    int d = 0;
    auto walk = [&]() {  // the failed lookup that precedes every insert: stop where a resident is closer to home
        const size_t mask = b_.size() - 1;
        for (i = hash20 & mask, d = 0; d <= b_[i].dist; ++d) i = (i + 1) & mask;
    };
    if (!b_.empty()) walk();
    // max_load_factor 0.5, growth x2 from 0 buckets; also the probe-length escape hatch (the 20-bit hash saturates)
    while (grow_next_ || d > kDistLimit || n_ >= threshold_) {
        // keys that agree on all 20 hash bits never separate: tsl would double until memory runs out; stop with an error instead
        if (b_.size() >= ((size_t)1 << 27)) throw ArgError("robin_map mirror: unbounded growth (more than 8192 voxels on one hash value)");
        std::vector<Bucket> bigger(b_.empty() ? 2 : b_.size() * 2);
        for (const auto &e : b_)
            if (e.dist >= 0) place(bigger, e, false);
        b_.swap(bigger);
        threshold_ = (size_t)((float)b_.size() * 0.5f);
        grow_next_ = false;
        walk();
    }
    const size_t mask = b_.size() - 1;
    Bucket e;
    e.hash = hash20, e.block = block, e.key = key, e.dist = d;
    if (b_[i].dist >= 0) {
        std::swap(e, b_[i]);
        // carry the evicted resident forward; it displaces anyone closer to home than itself
        for (d = e.dist + 1, i = (i + 1) & mask; b_[i].dist >= 0; ++d, i = (i + 1) & mask) {
            if (d <= b_[i].dist) continue;
            if (d >= kDistLimit) grow_next_ = true;
            e.dist = d;
            std::swap(e, b_[i]);
            d = e.dist;
        }
        e.dist = d;
    }
    b_[i] = e;
    ++n_;
}

void HostVoxelTable::erase_at(size_t i) {  // backward-shift deletion
    const size_t mask = b_.size() - 1;
    b_[i] = Bucket{};
    --n_;
    size_t prev = i, cur = (i + 1) & mask;
    while (b_[cur].dist > 0) {
        b_[prev] = b_[cur];
        b_[prev].dist -= 1;
        b_[cur] = Bucket{};
        prev = cur, cur = (cur + 1) & mask;
    }
}

uint32_t reference_voxel_hash(unsigned long long key) {  // core/VoxelHashMap.hpp:72-77
    int x, y, z;
    unpack_key(key, x, y, z);
    return ((1u << 20) - 1u) & ((uint32_t)x * 73856093u ^ (uint32_t)y * 19349663u ^ (uint32_t)z * 83492791u);
}

void VoxelMapGPU::set_eviction_faithful(bool on) {
    if (on == faithful_) return;
    if (!empty()) throw ArgError("the eviction mode can only be changed on an empty map");
    faithful_ = on;
    host_tbl_ = HostVoxelTable{};
    if (on && blk_cap_) {
        blk_first_.ensure(blk_cap_);
        blk_new_.ensure(blk_cap_);
        SAGE_CUDA(cudaMemsetAsync(blk_new_.p, 0, blk_new_.cap, stream_));
    }
}

// after a batch: the voxels it created, in the order the reference's sequential insert would have created them
void VoxelMapGPU::record_new_voxels(uint32_t m) {
    MapView v = view();
    new_list_.ensure((size_t)4 * m + 4);
    new_pin_.ensure((size_t)4 * m + 4);
    uint32_t *count = new_list_.p + (size_t)4 * m;
    SAGE_CUDA(cudaMemsetAsync(count, 0, sizeof(uint32_t), stream_));
    SAGE_LAUNCH(map_mark_first_kernel, blocks_for(m), kThreads, 0, stream_, v, upd_slot_.p, m);
    SAGE_LAUNCH(map_collect_new_kernel, blocks_for(m), kThreads, 0, stream_, v, upd_slot_.p, m, new_list_.p, count);
    SAGE_LAUNCH(map_clear_new_kernel, blocks_for(m), kThreads, 0, stream_, v, new_list_.p, count);
    SAGE_CUDA(cudaMemcpyAsync(new_pin_.p + (size_t)4 * m, count, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream_));
    SAGE_CUDA(cudaStreamSynchronize(stream_));
    const uint32_t k = new_pin_.p[(size_t)4 * m];
    if (k == 0) return;
    SAGE_CUDA(cudaMemcpyAsync(new_pin_.p, new_list_.p, (size_t)16 * k, cudaMemcpyDeviceToHost, stream_));
    SAGE_CUDA(cudaStreamSynchronize(stream_));
    std::vector<std::array<uint32_t, 4>> created(k);
    std::memcpy(created.data(), new_pin_.p, (size_t)16 * k);
    std::sort(created.begin(), created.end(), [](const auto &a, const auto &b) { return a[0] < b[0]; });
    for (const auto &c : created) {
        const unsigned long long key = (unsigned long long)c[2] | ((unsigned long long)c[3] << 32);
        host_tbl_.insert(key, reference_voxel_hash(key), c[1]);
    }
}

// the reference's sweep: range-for over the robin_map, erase(key) inside; after an erase the iterator moves on from the
// erased bucket, so the element shifted into it is not examined (core/VoxelHashMap.cpp:176-184)
void VoxelMapGPU::remove_far_faithful(double ox, double oy, double oz) {
    sync_stats();
    const uint32_t n_blocks = host_stats_.n_hi;
    if (n_blocks == 0) return;
    far_flags_.ensure(n_blocks);
    far_pin_.ensure(n_blocks);
    MapView v = view();
    SAGE_LAUNCH(map_far_flags_kernel, blocks_for(n_blocks), kThreads, 0, stream_, v, ox, oy, oz, max_distance_ * max_distance_, far_flags_.p, n_blocks);
    SAGE_CUDA(cudaMemcpyAsync(far_pin_.p, far_flags_.p, n_blocks, cudaMemcpyDeviceToHost, stream_));
    SAGE_CUDA(cudaStreamSynchronize(stream_));
    std::vector<uint32_t> gone;
    for (size_t i = 0; i < host_tbl_.bucket_count(); ++i) {
        const auto &e = host_tbl_.at(i);
        if (e.dist >= 0 && far_pin_.p[e.block]) {
            gone.push_back(e.block);
            host_tbl_.erase_at(i);
        }
    }
    if (gone.empty()) return;
    evict_list_.ensure(gone.size());
    evict_pin_.ensure(gone.size());
    std::memcpy(evict_pin_.p, gone.data(), gone.size() * sizeof(uint32_t));
    SAGE_CUDA(cudaMemcpyAsync(evict_list_.p, evict_pin_.p, gone.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream_));
    SAGE_LAUNCH(map_evict_list_kernel, blocks_for(gone.size()), kThreads, 0, stream_, v, evict_list_.p, (uint32_t)gone.size());
    SAGE_LAUNCH(tbl_clear_kernel, blocks_for(tbl_cap_), kThreads, 0, stream_, tbl_.p, tbl_cap_, ctrl_.p, 1);
    SAGE_LAUNCH(tbl_reinsert_kernel, blocks_for(n_blocks), kThreads, 0, stream_, v, 1);
    SAGE_LAUNCH(ctrl_finish_kernel, 1, 1, 0, stream_, ctrl_.p);
    SAGE_CUDA(cudaStreamSynchronize(stream_));  // evict_pin_ may be refilled by the next call
}

void VoxelMapGPU::remove_far(double ox, double oy, double oz) {
    set_device();
    if (hi_bound_ == 0) return;
    if (faithful_) return remove_far_faithful(ox, oy, oz);
    MapView v = view();
    SAGE_LAUNCH(map_evict_kernel, blocks_for(hi_bound_), kThreads, 0, stream_, v, ox, oy, oz, max_distance_ * max_distance_);
    SAGE_LAUNCH(tbl_clear_kernel, blocks_for(tbl_cap_), kThreads, 0, stream_, tbl_.p, tbl_cap_, ctrl_.p, 1);
    SAGE_LAUNCH(tbl_reinsert_kernel, blocks_for(hi_bound_), kThreads, 0, stream_, v, 1);
    SAGE_LAUNCH(ctrl_finish_kernel, 1, 1, 0, stream_, ctrl_.p);
}

long long VoxelMapGPU::dump(int32_t *keys, int32_t *counts, double *points, size_t cap_voxels) {
    set_device();
    sync_stats();
    const size_t hi = host_stats_.n_hi, live = host_stats_.n_live;
    if (!keys || cap_voxels < live) return (long long)live;
    std::vector<unsigned long long> k(hi);
    std::vector<int32_t> c(hi);
    std::vector<double> p(hi * (size_t)stride_ * 4);
    if (hi) {
        SAGE_CUDA(cudaMemcpyAsync(k.data(), blk_key_.p, hi * sizeof(unsigned long long), cudaMemcpyDeviceToHost, stream_));
        SAGE_CUDA(cudaMemcpyAsync(c.data(), blk_cnt_.p, hi * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
        SAGE_CUDA(cudaMemcpyAsync(p.data(), blk_pts_.p, hi * (size_t)stride_ * sizeof(double4), cudaMemcpyDeviceToHost, stream_));
        SAGE_CUDA(cudaStreamSynchronize(stream_));
    }
    size_t v = 0;
    // faithful mode: the reference's iteration order (bucket order of its robin_map); otherwise device block order
    std::vector<size_t> order;
    if (faithful_) {
        for (size_t i = 0; i < host_tbl_.bucket_count(); ++i)
            if (host_tbl_.at(i).dist >= 0) order.push_back(host_tbl_.at(i).block);
    } else {
        for (size_t b = 0; b < hi; ++b) order.push_back(b);
    }
    for (size_t b : order) {
        if (b >= hi || k[b] == kEmptyKey) continue;
        int x, y, z;
        unpack_key(k[b], x, y, z);
        keys[3 * v] = x, keys[3 * v + 1] = y, keys[3 * v + 2] = z;
        counts[v] = c[b];
        std::memset(points + v * (size_t)stride_ * 4, 0, sizeof(double) * 4 * (size_t)stride_);
        std::memcpy(points + v * (size_t)stride_ * 4, p.data() + b * (size_t)stride_ * 4, sizeof(double) * 4 * (size_t)c[b]);
        ++v;
    }
    return (long long)v;
}

long long VoxelMapGPU::pointcloud(double *out, size_t cap_points) {
    const long long total = num_points();
    if (!out || cap_points < (size_t)total) return total;
    const size_t live = host_stats_.n_live;
    std::vector<int32_t> keys(3 * live + 3), counts(live + 1);
    std::vector<double> pts((live + 1) * (size_t)stride_ * 4);
    const long long v = dump(keys.data(), counts.data(), pts.data(), live);
    size_t o = 0;
    for (long long b = 0; b < v; ++b) {
        std::memcpy(out + 4 * o, pts.data() + (size_t)b * stride_ * 4, sizeof(double) * 4 * (size_t)counts[b]);
        o += (size_t)counts[b];
    }
    return (long long)o;
}

void VoxelMapGPU::load(const int32_t *keys, const int32_t *counts, const double *points, int stride, size_t n_voxels) {
    set_device();
    clear();
    if (n_voxels == 0) return;
    reserve(n_voxels);
    std::vector<unsigned long long> k(n_voxels);
    std::vector<int32_t> c(n_voxels);
    std::vector<double> p(n_voxels * (size_t)stride_ * 4, 0.0);
    for (size_t v = 0; v < n_voxels; ++v) {
        if (!key_in_range(keys[3 * v], keys[3 * v + 1], keys[3 * v + 2])) throw ArgError("voxel key outside packable range");
        if (counts[v] < 0 || counts[v] > stride_ || counts[v] > stride) throw ArgError("voxel count exceeds basic+critical");
        k[v] = pack_key(keys[3 * v], keys[3 * v + 1], keys[3 * v + 2]);
        c[v] = counts[v];
        std::memcpy(p.data() + v * (size_t)stride_ * 4, points + v * (size_t)stride * 4, sizeof(double) * 4 * (size_t)counts[v]);
    }
    SAGE_CUDA(cudaMemcpyAsync(blk_key_.p, k.data(), n_voxels * sizeof(unsigned long long), cudaMemcpyHostToDevice, stream_));
    SAGE_CUDA(cudaMemcpyAsync(blk_cnt_.p, c.data(), n_voxels * sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
    SAGE_CUDA(cudaMemcpyAsync(blk_pts_.p, p.data(), p.size() * sizeof(double), cudaMemcpyHostToDevice, stream_));
    SAGE_CUDA(cudaMemsetAsync(blk_head_.p, 0xff, n_voxels * sizeof(uint32_t), stream_));
    SAGE_LAUNCH(ctrl_set_kernel, 1, 1, 0, stream_, ctrl_.p, (uint32_t)n_voxels, (uint32_t)n_voxels);
    hi_bound_ = live_bound_ = n_voxels;
    if (faithful_)
        for (size_t v = 0; v < n_voxels; ++v) host_tbl_.insert(k[v], reference_voxel_hash(k[v]), (uint32_t)v);
    SAGE_LAUNCH(map_build_hot_kernel, blocks_for(n_voxels * (size_t)stride_), kThreads, 0, stream_, view(), (uint32_t)n_voxels);
    rebuild_table(tbl_cap_);
    SAGE_CUDA(cudaStreamSynchronize(stream_));
}

void VoxelMapGPU::comm_init(int rank, int world, const uint8_t id[128]) {
    set_device();
    comm_destroy();
    comm_ = nccl_comm_create(rank, world, id);
}

// ---- fused all-reduce over NVLink peer memory ------------------------------------------------------------------
constexpr size_t kXchgDoubles = 2 * 8 * 24;  // [parity][source rank][slot], see registration.cu

void VoxelMapGPU::peer_handle(uint8_t out[64]) {
    set_device();
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (!xchg_local_) SAGE_CUDA(cudaMalloc(&xchg_local_, kXchgDoubles * sizeof(double)));
    // tags restart at 1 after every attach: clear stale ones now, before any peer can hold this buffer's handle
    SAGE_CUDA(cudaStreamSynchronize(stream_));
    SAGE_CUDA(cudaMemset(xchg_local_, 0, kXchgDoubles * sizeof(double)));
    SAGE_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t h;
    SAGE_CUDA(cudaIpcGetMemHandle(&h, xchg_local_));
    std::memcpy(out, &h, 64);
}

void VoxelMapGPU::peer_attach(int rank, int world, const uint8_t *handles) {
    set_device();
    if (world < 1 || world > 8 || rank < 0 || rank >= world) throw ArgError("peer_attach: 1..8 ranks of one node");
    if (!xchg_local_) throw ArgError("peer_attach: call sage_map_comm_peer_handle first");
    peer_detach();
    for (int k = 0; k < world; ++k) {
        if (k == rank) {
            peer_buf_[k] = xchg_local_;
            continue;
        }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, handles + 64 * (size_t)k, 64);
        void *ptr = nullptr;
        SAGE_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        peer_buf_[k] = (double *)ptr;
    }
    peer_rank_ = rank, peer_world_ = world, xchg_tag_ = 0, peer_inprocess_ = false;
}

double *VoxelMapGPU::peer_local_buffer() {
    set_device();
    if (!xchg_local_) SAGE_CUDA(cudaMalloc(&xchg_local_, kXchgDoubles * sizeof(double)));
    SAGE_CUDA(cudaStreamSynchronize(stream_));
    SAGE_CUDA(cudaMemset(xchg_local_, 0, kXchgDoubles * sizeof(double)));
    SAGE_CUDA(cudaDeviceSynchronize());
    return xchg_local_;
}

void VoxelMapGPU::peer_attach_local(int rank, int world, double *const *buffers, const int *devices) {
    set_device();
    if (world < 1 || world > 8 || rank < 0 || rank >= world) throw ArgError("peer_attach_local: 1..8 GPUs of one node");
    if (buffers[rank] != xchg_local_ || !xchg_local_) throw ArgError("peer_attach_local: buffers[rank] must be this map's own buffer");
    peer_detach();
    for (int k = 0; k < world; ++k) {
        if (k != rank) {
            int can = 0;
            SAGE_CUDA(cudaDeviceCanAccessPeer(&can, device_, devices[k]));
            if (!can) throw CudaError("GPU " + std::to_string(device_) + " cannot access GPU " + std::to_string(devices[k]) + " (no peer path): the sharded registration needs NVLink/PCIe peer access");
            const cudaError_t e = cudaDeviceEnablePeerAccess(devices[k], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled)
                (void)cudaGetLastError();
            else if (e != cudaSuccess)
                throw CudaError(std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        }
        peer_buf_[k] = buffers[k];
    }
    peer_rank_ = rank, peer_world_ = world, xchg_tag_ = 0, peer_inprocess_ = true;
}

void VoxelMapGPU::peer_detach() {
    for (int k = 0; k < 8; ++k) {
        if (!peer_inprocess_ && peer_buf_[k] && peer_buf_[k] != xchg_local_) cudaIpcCloseMemHandle(peer_buf_[k]);
        peer_buf_[k] = nullptr;
    }
    peer_world_ = 0;
    peer_inprocess_ = false;
}

void VoxelMapGPU::comm_destroy() {
    if (comm_) {
        nccl_comm_destroy(comm_);
        comm_ = nullptr;
    }
}

}  // namespace sage
