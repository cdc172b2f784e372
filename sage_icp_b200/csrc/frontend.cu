// Range crop + per-label-group voxel downsample (sm_100a) with the reference's output order.
//   Preprocess (range branch)   core/Preprocessing.cpp:173-187
//   VoxelDownsample             core/Preprocessing.cpp:44-84
//   DeSkewScan                  core/Deskew.cpp:36-50
// Which point survives ("first point by input order per (group, voxel)") is a pure set property, computed with a
// device hash grid and atomicMin on the point index.  The ORDER of the survivors is the iteration order of the
// reference's unreserved tsl::robin_map per group; it decides the ICP query set and
// the map insertion order, so it is reproduced exactly: the distinct keys' 20-bit hashes go to the host in
// first-index order, robin_iteration_order() replays the table, and the permutation comes back for the gather.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "frontend.cuh"
#include "device_sort.cuh"

namespace sage {

constexpr int kFeThreads = 256;
static inline unsigned fe_blocks(size_t n, int per_block = kFeThreads) { return (unsigned)((n + per_block - 1) / per_block); }

// ---------------------------------------------------------------------------------------------
// host: tsl::robin_map v1.0.1 iteration-order replay for distinct keys

void robin_iteration_order(const uint32_t *hash20, size_t n, uint32_t *order_out) {
    // One 64-bit word per bucket: [63:48] distance from the ideal bucket (0xffff = empty), [47:28] the key's 20-bit hash,
    // [27:0] payload (input position).  Two preallocated arrays alternate across growth generations.
    if (n >= (1u << 28)) throw ArgError("robin_iteration_order: too many keys");
    constexpr uint64_t kEmpty = ~0ull;
    constexpr int kDistLimit = 8192;
    static thread_local std::vector<uint64_t> buf[2], live_buf;
    size_t final_cap = 2;
    while (final_cap < 2 * n) final_cap *= 2;
    final_cap *= 2;  // a probe-length-forced growth (hash saturation) may add one generation
    for (auto &b : buf)
        if (b.size() < final_cap) b.resize(final_cap);
    if (live_buf.size() < n + 1) live_buf.resize(n + 1);
    uint64_t *cur = buf[0].data(), *nxt = buf[1].data();
    size_t B = 0, mask = 0, size = 0, load_threshold = 0;
    bool grow_next = false;
    auto dist_of = [](uint64_t e) { return e == kEmpty ? -1 : (int)(e >> 48); };
    auto make = [](int d, uint64_t rest) { return ((uint64_t)d << 48) | rest; };  // rest = hash << 28 | val

    // robin-hood swap-and-carry of `rest` from bucket ib with distance d
    // (tsl marks the table for growth when a CARRIED element is swapped in at a distance >= the limit; the new element's own
    // first swap is not checked)
    auto place = [&](uint64_t *T, size_t msk, size_t ib, int d, uint64_t rest, bool track) {
        bool carried = false;
        while (true) {
            const uint64_t e = T[ib];
            const int ed = dist_of(e);
            if (d > ed) {
                T[ib] = make(d, rest);
                if (ed < 0) return;
                if (track && carried && d >= kDistLimit) grow_next = true;
                carried = true;
                d = ed, rest = e & 0xffffffffffffull;
            }
            ++d;
            ib = (ib + 1) & msk;
        }
    };
    auto rehash = [&](size_t count) {
        if (count > final_cap) throw ArgError("robin_iteration_order: table growth beyond the preallocated generations");
        std::fill(nxt, nxt + count, kEmpty);
        const size_t nmask = count - 1;
        // The replay is bound by branch mispredictions, not by memory (every table fits the host's L2): whether a bucket of a
        // half-full table is occupied is a coin toss.  So the live entries are first compacted in bucket order WITHOUT a branch,
        // and an element whose new ideal bucket is still free — the common case right after doubling — is stored without entering
        // the swap-and-carry loop (1.5x on the whole function, 56 -> 38 ns per key on fresh keys).
        uint64_t *live = live_buf.data();
        size_t n_live = 0;
        for (size_t b = 0; b < B; ++b) {
            live[n_live] = cur[b];
            n_live += (cur[b] != kEmpty);
        }
        for (size_t j = 0; j < n_live; ++j) {
            const uint64_t rest = live[j] & 0xffffffffffffull;
            const size_t ib = (size_t)(rest >> 28) & nmask;
            if (nxt[ib] == kEmpty)
                nxt[ib] = rest;  // distance 0
            else
                place(nxt, nmask, ib, 0, rest, false);
        }
        std::swap(cur, nxt);
        B = count, mask = nmask;
        load_threshold = (size_t)((float)count * 0.5f);
    };

    for (size_t i = 0; i < n; ++i) {
        const uint32_t h = hash20[i] & 0xfffffu;
        size_t ib = 0;
        int d = 0;
        if (B) {
            ib = h & mask;
            while (d <= dist_of(cur[ib])) ib = (ib + 1) & mask, ++d;
        }
        while (grow_next || d > kDistLimit || size >= load_threshold) {
            rehash(B ? B * 2 : 2);
            grow_next = false;
            ib = h & mask, d = 0;
            while (d <= dist_of(cur[ib])) ib = (ib + 1) & mask, ++d;
        }
        place(cur, mask, ib, d, ((uint64_t)h << 28) | (uint64_t)i, true);
        ++size;
    }
    size_t k = 0;
    for (size_t b = 0; b < B; ++b)
        if (cur[b] != kEmpty) order_out[k++] = (uint32_t)(cur[b] & 0xfffffffu);  // (order_out holds exactly n entries: no branchless overrun)
}

// ---------------------------------------------------------------------------------------------
// kernels

constexpr int kDsBits = 20;
constexpr int kDsBias = 1 << (kDsBits - 1);

__device__ __forceinline__ bool crop_point(const CropParams &c, double4 &p) {
    if (!c.enabled) return true;
    const double nrm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(p.x, p.x), __dmul_rn(p.y, p.y)), __dmul_rn(p.z, p.z)));
    if (!(nrm < c.max_range && nrm > c.min_range)) return false;
    if (nrm > c.label_max_range) p.w = 0.0;
    return true;
}

__device__ __forceinline__ uint32_t reference_voxel_hash(int x, int y, int z) {  // core/Preprocessing.cpp:35-40
    return ((1u << 20) - 1u) & ((uint32_t)x * 73856093u ^ (uint32_t)y * 19349663u ^ (uint32_t)z * 83492791u);
}

__global__ void ds_insert_kernel(const double4 *in, uint32_t n, GroupTable g, CropParams crop, double scale,
                                 unsigned long long *tkey, uint32_t *tfirst, uint32_t mask, uint32_t *slot_out, uint32_t *err) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double4 p = in[i];
    slot_out[i] = kNil;
    if (!crop_point(crop, p)) return;
    const int label = __double2int_rz(p.w);
    int group = -1;
    for (int k = 0; k < g.n_labels; ++k)
        if (g.label[k] == label) {
            group = g.group_of[k];
            break;
        }
    if (group < 0) return;  // label in no group: dropped, core/Preprocessing.cpp:69
    const double s = __dmul_rn(g.voxel_size[group], scale);
    const int kx = trunc_div(p.x, s), ky = trunc_div(p.y, s), kz = trunc_div(p.z, s);
    if (kx <= -kDsBias || kx >= kDsBias || ky <= -kDsBias || ky >= kDsBias || kz <= -kDsBias || kz >= kDsBias) {
        atomicAdd(err, 1u);
        return;
    }
    const unsigned long long key = (unsigned long long)(uint32_t)(kx + kDsBias) | ((unsigned long long)(uint32_t)(ky + kDsBias) << kDsBits) |
                                   ((unsigned long long)(uint32_t)(kz + kDsBias) << (2 * kDsBits)) | ((unsigned long long)group << 60);
    uint32_t sl = (uint32_t)mix64(key) & mask;
    while (true) {
        const unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(tkey + sl);
        if (cur == key) break;
        if (cur == kEmptyKey) {
            const unsigned long long old = atomicCAS(tkey + sl, kEmptyKey, key);
            if (old == kEmptyKey || old == key) break;
        }
        sl = (sl + 1) & mask;
    }
    atomicMin(tfirst + sl, i);  // first point by input order wins, core/Preprocessing.cpp:71-72
    slot_out[i] = sl;
}

__global__ void ds_flag_kernel(const uint32_t *slot, const uint32_t *tfirst, uint32_t *flags, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t s = slot[i];
    flags[i] = (s != kNil && tfirst[s] == i) ? 1u : 0u;
}

__global__ void crop_flag_kernel(const double4 *in, uint32_t n, CropParams crop, uint32_t *flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double4 p = in[i];
    flags[i] = crop_point(crop, p) ? 1u : 0u;
}

__global__ void crop_scatter_kernel(const double4 *in, uint32_t n, CropParams crop, const uint32_t *flags, const uint32_t *pos, double4 *out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flags[i]) return;
    double4 p = in[i];
    crop_point(crop, p);
    out[pos[i]] = p;
}

__global__ void ds_collect_kernel(const uint32_t *slot, const uint32_t *flags, const uint32_t *pos, const unsigned long long *tkey,
                                  uint32_t *widx, uint32_t *whash, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flags[i]) return;
    const unsigned long long key = tkey[slot[i]];
    const unsigned long long m = (1ull << kDsBits) - 1;
    const int kx = (int)(key & m) - kDsBias, ky = (int)((key >> kDsBits) & m) - kDsBias, kz = (int)((key >> (2 * kDsBits)) & m) - kDsBias;
    const uint32_t group = (uint32_t)(key >> 60);
    const uint32_t j = pos[i];
    widx[j] = i;
    whash[j] = reference_voxel_hash(kx, ky, kz) | (group << 20);
}

__global__ void ds_gather_kernel(const double4 *in, const uint32_t *widx, const uint32_t *perm, CropParams crop, double4 *out, uint32_t m) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    double4 p = in[widx[perm[j]]];
    crop_point(crop, p);  // re-applies the far-label zeroing (core/Preprocessing.cpp:178)
    out[j] = p;
}

// exclusive scan of 0/1 flags, 1024 elements per block
__global__ void scan_block_kernel(const uint32_t *flags, uint32_t *pos, uint32_t *block_sums, uint32_t n) {
    __shared__ uint32_t s[kFeThreads];
    const uint32_t base = blockIdx.x * 1024u + threadIdx.x * 4u;
    uint32_t v[4], e[4], run = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k] = (base + k < n) ? flags[base + k] : 0u;
        e[k] = run;
        run += v[k];
    }
    s[threadIdx.x] = run;
    __syncthreads();
    for (int o = 1; o < kFeThreads; o <<= 1) {
        const uint32_t t = threadIdx.x >= (unsigned)o ? s[threadIdx.x - o] : 0u;
        __syncthreads();
        s[threadIdx.x] += t;
        __syncthreads();
    }
    const uint32_t excl = s[threadIdx.x] - run;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (base + k < n) pos[base + k] = excl + e[k];
    if (threadIdx.x == kFeThreads - 1) block_sums[blockIdx.x] = s[threadIdx.x];
}
__global__ void scan_sums_kernel(uint32_t *block_sums, uint32_t nb, uint32_t *total, const uint32_t *err) {
    if (err) total[1] = *err;
    uint32_t run = 0;
    for (uint32_t b = 0; b < nb; ++b) {
        const uint32_t v = block_sums[b];
        block_sums[b] = run;
        run += v;
    }
    total[0] = run;
}
__global__ void scan_add_kernel(uint32_t *pos, const uint32_t *block_sums, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pos[i] += block_sums[i / 1024u];
}

// ---------------------------------------------------------------------------------------------
// Dynamic-vehicle filter (core/Preprocessing.cpp:95-172).  PCL's Euclidean cluster extraction is single linkage with
// "squared float distance < tolerance^2", i.e. connected components; they are found with a 0.5 m cell grid and a lock-free
// union-find whose roots are the smallest member index (so the result does not depend on scheduling).
constexpr uint32_t kClsDrop = 0, kClsInlier = 1, kClsVehicle = 2, kClsLandmarkBit = 4;

__device__ __forceinline__ bool dyn_near(float ax, float ay, float az, float bx, float by, float bz) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)) < 0.25f;  // 0.5 m, FLANN keeps dist < radius
}
__device__ __forceinline__ unsigned long long dyn_cell_key(int cx, int cy, int cz) {
    return (unsigned long long)(uint32_t)(cx + kKeyBias) | ((unsigned long long)(uint32_t)(cy + kKeyBias) << kKeyBits) |
           ((unsigned long long)(uint32_t)(cz + kKeyBias) << (2 * kKeyBits));
}
__device__ __forceinline__ int dyn_cell(float v) { return __float2int_rd(__fmul_rn(v, 2.0f)); }  // 0.5 m cells

__device__ __forceinline__ uint32_t dyn_find_slot(const unsigned long long *keys, uint32_t mask, unsigned long long key) {
    uint32_t sl = (uint32_t)mix64(key) & mask;
    while (true) {
        const unsigned long long cur = keys[sl];
        if (cur == key) return sl;
        if (cur == kEmptyKey) return kNil;
        sl = (sl + 1) & mask;
    }
}

// classify + put vehicle / landmark points on their cell's list
// Vehicle points go on a FINE grid (0.288 m cells: any two points of one cell are closer than 0.288 * sqrt(3) = 0.4988 m, i.e.
// connected by definition), landmark points on a 0.5 m grid (the radius of the hit count).  Both grids share one key table;
// bit 63 of the key tells them apart.
constexpr float kFineInv = 1.0f / 0.288f;
__device__ __forceinline__ int dyn_fine_cell(float v) { return __float2int_rd(__fmul_rn(v, kFineInv)); }
__device__ __forceinline__ uint32_t dyn_insert_slot(unsigned long long *keys, uint32_t mask, unsigned long long key) {
    uint32_t sl = (uint32_t)mix64(key) & mask;
    while (true) {
        const unsigned long long cur = *reinterpret_cast<volatile unsigned long long *>(keys + sl);
        if (cur == key) return sl;
        if (cur == kEmptyKey) {
            const unsigned long long old = atomicCAS(keys + sl, kEmptyKey, key);
            if (old == kEmptyKey || old == key) return sl;
        }
        sl = (sl + 1) & mask;
    }
}

__global__ void dyn_classify_kernel(const double4 *in, uint32_t n, CropParams crop, DynFilterParams dp, unsigned long long *keys, uint32_t mask,
                                    uint32_t *head_v, uint32_t *head_l, uint32_t *next_v, uint32_t *next_l, uint32_t *parent, uint32_t *cls,
                                    uint32_t *err) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double4 p = in[i];
    parent[i] = i;
    if (!crop_point(crop, p)) {
        cls[i] = kClsDrop;
        return;
    }
    // point_temp.label = static_cast<uint32_t>(point_new[3]); membership tests on that integer (:108,:110,:155)
    const int label = (int)(uint32_t)__double2uint_rz(p.w);
    bool vehicle = false, landmark = false;
    for (int k = 0; k < dp.n_dynamic; ++k) vehicle |= dp.dynamic_labels[k] == label;
    for (int k = 0; k < dp.n_landmark; ++k) landmark |= dp.landmark_labels[k] == label;
    cls[i] = (vehicle ? kClsVehicle : kClsInlier) | (landmark ? kClsLandmarkBit : 0u);
    if (!vehicle && !landmark) return;
    const float x = __double2float_rn(p.x), y = __double2float_rn(p.y), z = __double2float_rn(p.z);
    const int fx = dyn_fine_cell(x), fy = dyn_fine_cell(y), fz = dyn_fine_cell(z);
    const int cx = dyn_cell(x), cy = dyn_cell(y), cz = dyn_cell(z);
    if (!key_in_range(fx, fy, fz) || !key_in_range(cx, cy, cz)) {
        atomicAdd(err, 1u);
        return;
    }
    if (vehicle) {
        const uint32_t sl = dyn_insert_slot(keys, mask, dyn_cell_key(fx, fy, fz) | (1ull << 63));
        next_v[i] = atomicExch(head_v + sl, i);
    }
    if (landmark) {
        const uint32_t sl = dyn_insert_slot(keys, mask, dyn_cell_key(cx, cy, cz));
        next_l[i] = atomicExch(head_l + sl, i);
    }
}

__device__ __forceinline__ uint32_t dyn_root(const uint32_t *parent, uint32_t a) {
    while (true) {
        const uint32_t pa = *reinterpret_cast<const volatile uint32_t *>(parent + a);
        if (pa == a) return a;
        a = pa;
    }
}

__device__ __forceinline__ void dyn_unite(uint32_t *parent, uint32_t a, uint32_t b) {
    while (true) {
        a = dyn_root(parent, a), b = dyn_root(parent, b);
        if (a == b) return;
        if (a < b) {
            const uint32_t t = a;
            a = b, b = t;
        }
        if (atomicCAS(parent + a, a, b) == a) return;  // hang the larger root under the smaller: roots end as minimum indices
    }
}

// Single-linkage clusters of the vehicle points (= PCL's EuclideanClusterExtraction with tolerance 0.5).  A point joins its own
// fine cell (all of it is within 0.5 m), then for each of the 124 fine cells that can hold a point within 0.5 m it looks for ONE
// such point — cell mates of that point are already tied to it — and skips cells that already belong to its component.
__global__ void dyn_union_kernel(const double4 *in, uint32_t n, const unsigned long long *keys, uint32_t mask, const uint32_t *head_v,
                                 const uint32_t *next_v, uint32_t *parent, const uint32_t *cls) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (cls[i] & 3u) != kClsVehicle) return;
    const double4 p = in[i];
    const float x = __double2float_rn(p.x), y = __double2float_rn(p.y), z = __double2float_rn(p.z);
    const int fx = dyn_fine_cell(x), fy = dyn_fine_cell(y), fz = dyn_fine_cell(z);
    for (int c = 0; c < 125; ++c) {
        const int nx = fx + c / 25 - 2, ny = fy + (c / 5) % 5 - 2, nz = fz + c % 5 - 2;
        if (!key_in_range(nx, ny, nz)) continue;
        const uint32_t sl = dyn_find_slot(keys, mask, dyn_cell_key(nx, ny, nz) | (1ull << 63));
        if (sl == kNil) continue;
        const uint32_t first = head_v[sl];
        if (c == 62) {  // own cell
            if (first != i) dyn_unite(parent, i, first);
            continue;
        }
        if (dyn_root(parent, first) == dyn_root(parent, i)) continue;  // already one component (cell mates share a root eventually)
        for (uint32_t j = first; j != kNil; j = next_v[j]) {
            const double4 q = in[j];
            if (dyn_near(x, y, z, __double2float_rn(q.x), __double2float_rn(q.y), __double2float_rn(q.z))) {
                dyn_unite(parent, i, j);
                break;
            }
        }
    }
}

// cluster sizes
__global__ void dyn_size_kernel(uint32_t n, const uint32_t *parent, const uint32_t *cls, uint32_t *csize) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (cls[i] & 3u) != kClsVehicle) return;
    atomicAdd(csize + dyn_root(parent, i), 1u);
}

// per vehicle point: final root, cluster size, landmark hits within 0.5 m (summed per cluster, as the reference's count_size)
__global__ void dyn_count_kernel(const double4 *in, uint32_t n, const unsigned long long *keys, uint32_t mask, const uint32_t *head_l,
                                 const uint32_t *next_l, uint32_t *parent, const uint32_t *cls, const uint32_t *csize, uint32_t *clm, double dy_th) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || (cls[i] & 3u) != kClsVehicle) return;
    const uint32_t r = dyn_root(parent, i);
    // the verdict is "hits > int(dy_th * size)" (:159): clusters too small to be kept, or already over the bar, need no more counting
    // (the reference breaks out of its loops the same way)
    const uint32_t size = csize[r];
    if (size < 5u || (long long)*reinterpret_cast<volatile uint32_t *>(clm + r) > (long long)__double2int_rz(__dmul_rn(dy_th, (double)size))) return;
    const double4 p = in[i];
    const float x = __double2float_rn(p.x), y = __double2float_rn(p.y), z = __double2float_rn(p.z);
    const int cx = dyn_cell(x), cy = dyn_cell(y), cz = dyn_cell(z);
    uint32_t hits = 0;
    for (int c = 0; c < 27; ++c) {
        const int nx = cx + c / 9 - 1, ny = cy + (c / 3) % 3 - 1, nz = cz + c % 3 - 1;
        if (!key_in_range(nx, ny, nz)) continue;
        const uint32_t sl = dyn_find_slot(keys, mask, dyn_cell_key(nx, ny, nz));
        if (sl == kNil) continue;
        for (uint32_t j = head_l[sl]; j != kNil; j = next_l[j]) {
            const double4 q = in[j];
            hits += dyn_near(x, y, z, __double2float_rn(q.x), __double2float_rn(q.y), __double2float_rn(q.z)) ? 1u : 0u;
        }
    }
    if (hits) atomicAdd(clm + r, hits);
}

// flags[i] = plain inlier, flags[n + i] = vehicle point of a kept cluster: one scan orders inliers first, vehicles after
__global__ void dyn_flag_kernel(uint32_t n, const uint32_t *parent, const uint32_t *cls, const uint32_t *csize, const uint32_t *clm, double dy_th,
                                uint32_t *flags) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t c = cls[i] & 3u;
    uint32_t keep_vehicle = 0;
    if (c == kClsVehicle) {
        const uint32_t r = dyn_root(parent, i);
        const uint32_t size = csize[r];
        // setMinClusterSize(5) (:137); is_static_vehicle iff count_size > int(dy_th * cluster_size) (:159)
        keep_vehicle = (size >= 5u && (long long)clm[r] > (long long)__double2int_rz(__dmul_rn(dy_th, (double)size))) ? 1u : 0u;
    }
    flags[i] = c == kClsInlier ? 1u : 0u;
    flags[n + i] = keep_vehicle;
}

__global__ void dyn_scatter_kernel(const double4 *in, uint32_t n, CropParams crop, const uint32_t *flags, const uint32_t *pos, double4 *out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n || !flags[t]) return;  // the plain inliers, input order
    double4 p = in[t];
    crop_point(crop, p);  // re-applies the far-label zeroing
    out[pos[t]] = p;
}

// The reference appends the kept vehicle points cluster by cluster (core/Preprocessing.cpp:141-170) in the order PCL's
// EuclideanClusterExtraction returns the clusters: by descending size, indices inside a cluster ascending (extractEuclideanClusters
// sorts each cluster's indices, then sorts the clusters by size).  Equal sizes: discovery order, i.e. ascending smallest member —
// which is this union-find's root.  Sort key of a kept vehicle point: (n - size, root); a stable sort keeps the index order inside a
// cluster.  Points that are not kept get the bit above the key (they sort to the end and are not emitted).
__global__ void dyn_sortkey_kernel(uint32_t n, const uint32_t *parent, const uint32_t *flags, const uint32_t *csize, int nbits, int end_bit,
                                   unsigned long long *keys, uint32_t *vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned long long k = 1ull << end_bit;
    if (flags[n + i]) {
        const uint32_t r = dyn_root(parent, i);
        k = ((unsigned long long)(n - csize[r]) << nbits) | r;
    }
    keys[i] = k, vals[i] = i;
}
__global__ void dyn_scatter_vehicle_kernel(const double4 *in, uint32_t n, CropParams crop, const unsigned long long *keys, const uint32_t *vals,
                                           int end_bit, const uint32_t *pos, double4 *out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n || (keys[j] >> end_bit)) return;
    double4 p = in[vals[j]];
    crop_point(crop, p);
    out[pos[n] + j] = p;  // pos[n] = number of plain inliers (exclusive scan over [inlier flags | vehicle flags])
}

// one thread per point; fields may sit at any byte offset (the reference's message is a packed 17-byte record:
// f32 x, y, z @0/4/8, u8 label @12, u32 rgb @13 — eval/kitti_pub.py:184-207), so they are assembled from bytes
__device__ __forceinline__ float load_f32_unaligned(const uint8_t *p) {
    const uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
    return __uint_as_float(v);
}
__global__ void unpack_pointcloud2_kernel(const uint8_t *data, uint32_t n, uint32_t step, uint32_t xo, uint32_t yo, uint32_t zo, uint32_t lo,
                                          int label_is_f32, double4 *out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *r = data + (size_t)i * step;
    // points.emplace_back(*msg_x, *msg_y, *msg_z, *msg_l): f32 / u8 widened to f64 (ros/ros2/Utils.hpp:170,176)
    const double l = label_is_f32 ? (double)load_f32_unaligned(r + lo) : (double)r[lo];
    out[i] = make_double4((double)load_f32_unaligned(r + xo), (double)load_f32_unaligned(r + yo), (double)load_f32_unaligned(r + zo), l);
}

__global__ void deskew_kernel(const double4 *in, const double *ts, uint32_t n, double d0, double d1, double d2, double d3, double d4,
                              double d5, double4 *out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double s = ts[i] - 0.5;  // mid_pose_timestamp, core/Deskew.cpp:33
    const double xi[6] = {s * d0, s * d1, s * d2, s * d3, s * d4, s * d5};
    const Pose motion = pose_exp(xi);
    const double4 p = in[i];
    double x, y, z;
    pose_act(motion, p.x, p.y, p.z, x, y, z);
    out[i] = make_double4(x, y, z, p.w);
}

// ---------------------------------------------------------------------------------------------
// host

FrontEnd::FrontEnd(const GroupTable &groups, int device, cudaStream_t stream) : groups_(groups), device_(device), stream_(stream) {
    SAGE_CUDA(cudaSetDevice(device_));
    preload_frontend_kernels();
    total_.ensure(2);
    total_pin_.ensure(2);
    if (const char *e = getenv("SAGE_FE_TRACE")) trace_ = atoi(e) != 0;
}

FrontEnd::~FrontEnd() {
    if (trace_ && n_calls_ > 0)
        fprintf(stderr,
                "[sage front end] %lld VoxelDownsample calls, %.1f distinct keys per call; host us per call: enqueue %.1f, wait for the device %.1f, "
                "group split %.1f, robin replay %.1f, permutation + gather launch %.1f\n",
                n_calls_, (double)n_keys_ / n_calls_, 1e6 * t_enqueue_ / n_calls_, 1e6 * t_wait_ / n_calls_, 1e6 * t_group_ / n_calls_,
                1e6 * t_replay_ / n_calls_, 1e6 * t_tail_ / n_calls_);
}

void FrontEnd::scan_flags(size_t n, uint32_t *total_out, const uint32_t *err) {
    const unsigned nb = fe_blocks(n, 1024);
    block_sums_.ensure(nb);
    pos_.ensure(n);
    SAGE_LAUNCH(scan_block_kernel, nb, kFeThreads, 0, stream_, flags_.p, pos_.p, block_sums_.p, (uint32_t)n);
    SAGE_LAUNCH(scan_sums_kernel, 1, 1, 0, stream_, block_sums_.p, nb, total_out, err);
    SAGE_LAUNCH(scan_add_kernel, fe_blocks(n), kFeThreads, 0, stream_, pos_.p, block_sums_.p, (uint32_t)n);
}

size_t FrontEnd::preprocess(const double4 *in, size_t n, const CropParams &crop, double4 *out) {
    SAGE_CUDA(cudaSetDevice(device_));
    if (n == 0) return 0;
    flags_.ensure(n);
    SAGE_LAUNCH(crop_flag_kernel, fe_blocks(n), kFeThreads, 0, stream_, in, (uint32_t)n, crop, flags_.p);
    scan_flags(n, total_.p, nullptr);
    SAGE_LAUNCH(crop_scatter_kernel, fe_blocks(n), kFeThreads, 0, stream_, in, (uint32_t)n, crop, flags_.p, pos_.p, out);
    SAGE_CUDA(cudaMemcpyAsync(total_pin_.p, total_.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream_));
    SAGE_CUDA(cudaStreamSynchronize(stream_));
    return total_pin_.p[0];
}

size_t FrontEnd::downsample(const double4 *in, size_t n, double vox_scale, const CropParams &crop, double4 *out) {
    SAGE_CUDA(cudaSetDevice(device_));
    if (n == 0) return 0;
    uint32_t cap = 1024;
    while ((size_t)cap < 2 * n) cap *= 2;
    if (cap > tcap_) {
        tkey_.ensure(cap);
        tfirst_.ensure(cap);
        tcap_ = cap;
    }
    slot_.ensure(n);
    flags_.ensure(n);
    widx_.ensure(n);
    // Host-visible results live in pinned memory that the kernels write / read directly (zero copy), so one stream
    // synchronisation per call is enough: [count, range errors] + the survivors' hashes out, the permutation back in.
    // Two permutation buffers alternate between calls: the gather of call k may still be running when call k+1 fills its own.
    whash_pin_.ensure(n);
    perm_pin_[parity_].ensure(n);  // this buffer's last reader (two calls ago) finished before the previous call's synchronisation
    uint32_t *perm_host = perm_pin_[parity_].p;
    parity_ ^= 1;
    using fe_clock = std::chrono::steady_clock;
    const auto tr0 = fe_clock::now();
    SAGE_CUDA(cudaMemsetAsync(total_.p, 0, 2 * sizeof(uint32_t), stream_));
    SAGE_CUDA(cudaMemsetAsync(tkey_.p, 0xff, (size_t)cap * sizeof(unsigned long long), stream_));
    SAGE_CUDA(cudaMemsetAsync(tfirst_.p, 0xff, (size_t)cap * sizeof(uint32_t), stream_));
    SAGE_LAUNCH(ds_insert_kernel, fe_blocks(n), kFeThreads, 0, stream_, in, (uint32_t)n, groups_, crop, vox_scale, tkey_.p, tfirst_.p,
                cap - 1, slot_.p, total_.p + 1);
    SAGE_LAUNCH(ds_flag_kernel, fe_blocks(n), kFeThreads, 0, stream_, slot_.p, tfirst_.p, flags_.p, (uint32_t)n);
    scan_flags(n, total_pin_.p, total_.p + 1);  // count and range-error flag land in pinned memory
    SAGE_LAUNCH(ds_collect_kernel, fe_blocks(n), kFeThreads, 0, stream_, slot_.p, flags_.p, pos_.p, tkey_.p, widx_.p, whash_pin_.p, (uint32_t)n);
    const auto tr1 = fe_clock::now();
    SAGE_CUDA(cudaStreamSynchronize(stream_));
    const auto tr2 = fe_clock::now();
    if (total_pin_.p[1]) throw ArgError("VoxelDownsample: point/voxel_size outside the +-2^19 voxel range");
    const size_t m = total_pin_.p[0];
    if (m == 0) return 0;

    // per group: replay the robin_map on the distinct keys (first-index order) and emit groups in index order
    // (core/Preprocessing.cpp:76-82).  Groups are independent tables: they are replayed side by side on host threads.
    const int G = groups_.n_groups;
    group_members_.resize(G);
    group_hashes_.resize(G);
    group_order_.resize(G);
    for (int g = 0; g < G; ++g) group_members_[g].clear(), group_hashes_[g].clear();
    for (size_t j = 0; j < m; ++j) {
        const uint32_t w = whash_pin_.p[j], g = w >> 20;
        group_members_[g].push_back((uint32_t)j);
        group_hashes_[g].push_back(w & 0xfffffu);
    }
    const auto tr3 = fe_clock::now();
#pragma omp parallel for schedule(dynamic, 1) num_threads(G < 6 ? G : 6) if (m > 1500)
    for (int g = 0; g < G; ++g) {
        group_order_[g].resize(group_members_[g].size());
        if (!group_members_[g].empty())
            robin_iteration_order(group_hashes_[g].data(), group_hashes_[g].size(), group_order_[g].data());
    }
    const auto tr4 = fe_clock::now();
    size_t o = 0;
    for (int g = 0; g < G; ++g)
        for (size_t k = 0; k < group_order_[g].size(); ++k) perm_host[o++] = group_members_[g][group_order_[g][k]];
    SAGE_LAUNCH(ds_gather_kernel, fe_blocks(m), kFeThreads, 0, stream_, in, widx_.p, perm_host, crop, out, (uint32_t)m);
    if (trace_) {
        const auto tr5 = fe_clock::now();
        auto sec = [](fe_clock::time_point a, fe_clock::time_point b) { return std::chrono::duration<double>(b - a).count(); };
        t_enqueue_ += sec(tr0, tr1), t_wait_ += sec(tr1, tr2), t_group_ += sec(tr2, tr3), t_replay_ += sec(tr3, tr4), t_tail_ += sec(tr4, tr5);
        n_calls_ += 1, n_keys_ += (long long)m;
    }
    return m;  // `out` is valid in stream order; callers that read it on the host synchronise themselves
}

size_t FrontEnd::preprocess_dynamic(const double4 *in, size_t n, const CropParams &crop, const DynFilterParams &dyn, double4 *out) {
    SAGE_CUDA(cudaSetDevice(device_));
    if (n == 0) return 0;
    uint32_t cap = 1024;
    while ((size_t)cap < 4 * n) cap *= 2;  // two grids share the table
    if (cap > cell_cap_) {
        cell_key_.ensure(cap);
        cell_head_v_.ensure(cap);
        cell_head_l_.ensure(cap);
        cell_cap_ = cap;
    }
    next_v_.ensure(n), next_l_.ensure(n), parent_.ensure(n), csize_.ensure(n), clm_.ensure(n), cls_.ensure(n);
    flags_.ensure(2 * n);
    SAGE_CUDA(cudaMemsetAsync(cell_key_.p, 0xff, (size_t)cap * sizeof(unsigned long long), stream_));
    SAGE_CUDA(cudaMemsetAsync(cell_head_v_.p, 0xff, (size_t)cap * sizeof(uint32_t), stream_));
    SAGE_CUDA(cudaMemsetAsync(cell_head_l_.p, 0xff, (size_t)cap * sizeof(uint32_t), stream_));
    SAGE_CUDA(cudaMemsetAsync(csize_.p, 0, n * sizeof(uint32_t), stream_));
    SAGE_CUDA(cudaMemsetAsync(clm_.p, 0, n * sizeof(uint32_t), stream_));
    SAGE_CUDA(cudaMemsetAsync(total_.p, 0, 2 * sizeof(uint32_t), stream_));
    const uint32_t nn = (uint32_t)n;
    SAGE_LAUNCH(dyn_classify_kernel, fe_blocks(n), kFeThreads, 0, stream_, in, nn, crop, dyn, cell_key_.p, cap - 1, cell_head_v_.p, cell_head_l_.p,
                next_v_.p, next_l_.p, parent_.p, cls_.p, total_.p + 1);
    SAGE_LAUNCH(dyn_union_kernel, fe_blocks(n), kFeThreads, 0, stream_, in, nn, cell_key_.p, cap - 1, cell_head_v_.p, next_v_.p, parent_.p, cls_.p);
    SAGE_LAUNCH(dyn_size_kernel, fe_blocks(n), kFeThreads, 0, stream_, nn, parent_.p, cls_.p, csize_.p);
    SAGE_LAUNCH(dyn_count_kernel, fe_blocks(n), kFeThreads, 0, stream_, in, nn, cell_key_.p, cap - 1, cell_head_l_.p, next_l_.p, parent_.p, cls_.p,
                csize_.p, clm_.p, dyn.dy_th);
    SAGE_LAUNCH(dyn_flag_kernel, fe_blocks(n), kFeThreads, 0, stream_, nn, parent_.p, cls_.p, csize_.p, clm_.p, dyn.dy_th, flags_.p);
    scan_flags(2 * n, total_pin_.p, total_.p + 1);
    SAGE_LAUNCH(dyn_scatter_kernel, fe_blocks(n), kFeThreads, 0, stream_, in, nn, crop, flags_.p, pos_.p, out);
    {
        int nbits = 1;
        while (((size_t)1 << nbits) < n) ++nbits;  // roots < n <= 2^nbits; n - size <= n < 2^(nbits + 1)
        const int end_bit = 2 * nbits + 1;
        dyn_key_[0].ensure(n), dyn_key_[1].ensure(n), dyn_val_[0].ensure(n), dyn_val_[1].ensure(n);
        const size_t tmp = sort_pairs_tmp_bytes_u64(n, end_bit + 1);
        sort_tmp_.ensure(tmp ? tmp : 1);
        SAGE_LAUNCH(dyn_sortkey_kernel, fe_blocks(n), kFeThreads, 0, stream_, nn, parent_.p, flags_.p, csize_.p, nbits, end_bit, dyn_key_[0].p,
                    dyn_val_[0].p);
        g_launches.fetch_add(sort_pairs_u64(sort_tmp_.p, tmp, dyn_key_[0].p, dyn_key_[1].p, dyn_val_[0].p, dyn_val_[1].p, n, end_bit + 1, stream_),
                             std::memory_order_relaxed);
        SAGE_LAUNCH(dyn_scatter_vehicle_kernel, fe_blocks(n), kFeThreads, 0, stream_, in, nn, crop, dyn_key_[1].p, dyn_val_[1].p, end_bit, pos_.p, out);
    }
    SAGE_CUDA(cudaStreamSynchronize(stream_));
    if (total_pin_.p[1]) throw ArgError("Preprocess: point outside the +-2^20 cell range of the dynamic-vehicle filter");
    return total_pin_.p[0];
}

void FrontEnd::unpack_pointcloud2(const uint8_t *data_dev, size_t n, uint32_t point_step, uint32_t x_off, uint32_t y_off, uint32_t z_off,
                                  uint32_t label_off, int label_is_f32, double4 *out) {
    SAGE_CUDA(cudaSetDevice(device_));
    if (n == 0) return;
    SAGE_LAUNCH(unpack_pointcloud2_kernel, fe_blocks(n), kFeThreads, 0, stream_, data_dev, (uint32_t)n, point_step, x_off, y_off, z_off, label_off,
                label_is_f32, out);
}

// Key-frame occupancy grid of the ROS node (utils::EigenToGridMap, ros/ros2/Utils.hpp:220-242), optionally of the points moved into
// the last key frame first (sageICP::TransformToLastFrame -> TransformPoints, pipeline/sageICP.cpp:123-129), and the overlap ratio
// against a previous grid (utils::compute_occ_overlap, Utils.hpp:244-258).  One thread per point; cells are 0/1, so plain stores race
// benignly; the two counts are integer atomics (exact, order-independent).
__global__ void occ_grid_kernel(const double4 *pts, uint32_t n, Pose T, int apply, OccGridParams g, int32_t *grid) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double4 p = pts[i];
    double x = p.x, y = p.y, z = p.z;
    if (apply) pose_act(T, p.x, p.y, p.z, x, y, z);
    // if (point[k] < lo || point[k] > hi) continue;  (Utils.hpp:227-231; NaN compares false on both sides and goes on, as there)
    if (x < g.x0 || x > g.x1 || y < g.y0 || y > g.y1 || z < g.z0 || z > g.z1) return;
    // int occ_x = static_cast<int>((point[0] + bounds[0][1]) / x_resolution);  (:233-234: the UPPER bound is added, as written)
    const double fx = __ddiv_rn(__dadd_rn(x, g.x1), g.x_res), fy = __ddiv_rn(__dadd_rn(y, g.y1), g.y_res);
    if (!(fx > -2147483648.0 && fx < 2147483648.0 && fy > -2147483648.0 && fy < 2147483648.0)) return;  // int conversion would be UB there
    const int ox = __double2int_rz(fx), oy = __double2int_rz(fy);
    if (ox >= 0 && ox < g.cols && oy >= 0 && oy < g.rows) grid[(size_t)oy * g.cols + ox] = 1;
}
__global__ void occ_overlap_kernel(const int32_t *occ_s, const int32_t *occ_t, uint32_t cells, unsigned *counts) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool s1 = i < cells && occ_s[i] == 1;
    const unsigned both = __ballot_sync(0xffffffffu, s1 && occ_t[i] == 1), tot = __ballot_sync(0xffffffffu, s1);
    if ((threadIdx.x & 31) == 0) {
        if (both) atomicAdd(counts, (unsigned)__popc(both));
        if (tot) atomicAdd(counts + 1, (unsigned)__popc(tot));
    }
}

void FrontEnd::key_frame_grid(const double4 *pts, size_t n, const Pose *T, const OccGridParams &g, const int32_t *last_occ_host, int32_t *grid_host,
                              double *overlap) {
    SAGE_CUDA(cudaSetDevice(device_));
    const size_t cells = (size_t)g.rows * g.cols;
    occ_.ensure(2 * cells + 2);
    int32_t *cur = reinterpret_cast<int32_t *>(occ_.p), *last = cur + cells;
    unsigned *counts = reinterpret_cast<unsigned *>(occ_.p + 2 * cells);
    SAGE_CUDA(cudaMemsetAsync(cur, 0, cells * sizeof(int32_t), stream_));
    if (n)
        SAGE_LAUNCH(occ_grid_kernel, fe_blocks(n), kFeThreads, 0, stream_, pts, (uint32_t)n, T ? *T : pose_identity(), T ? 1 : 0, g, cur);
    if (last_occ_host) {
        SAGE_CUDA(cudaMemcpyAsync(last, last_occ_host, cells * sizeof(int32_t), cudaMemcpyHostToDevice, stream_));
        SAGE_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(unsigned), stream_));
        SAGE_LAUNCH(occ_overlap_kernel, fe_blocks(cells), kFeThreads, 0, stream_, last, cur, (uint32_t)cells, counts);
    }
    unsigned c[2] = {0, 0};
    if (grid_host) SAGE_CUDA(cudaMemcpyAsync(grid_host, cur, cells * sizeof(int32_t), cudaMemcpyDeviceToHost, stream_));
    if (last_occ_host) SAGE_CUDA(cudaMemcpyAsync(c, counts, sizeof(c), cudaMemcpyDeviceToHost, stream_));
    SAGE_CUDA(cudaStreamSynchronize(stream_));
    // return static_cast<double>(overlap) / total;  (Utils.hpp:257: 0 / 0 = NaN when the last grid is empty, as there)
    if (overlap) *overlap = last_occ_host ? (double)c[0] / (double)c[1] : 0.0;
}

void FrontEnd::deskew(const double4 *in, const double *ts, size_t n, const Pose &start, const Pose &finish, double4 *out) {
    SAGE_CUDA(cudaSetDevice(device_));
    if (n == 0) return;
    double d[6];
    pose_log(pose_mul(pose_inverse(start), finish), d);  // core/Deskew.cpp:40
    SAGE_LAUNCH(deskew_kernel, fe_blocks(n), kFeThreads, 0, stream_, in, ts, (uint32_t)n, d[0], d[1], d[2], d[3], d[4], d[5], out);
}

void preload_frontend_kernels() {
    const void *ks[] = {(const void *)ds_insert_kernel, (const void *)ds_flag_kernel, (const void *)crop_flag_kernel, (const void *)crop_scatter_kernel, (const void *)ds_collect_kernel, (const void *)ds_gather_kernel, (const void *)scan_block_kernel, (const void *)scan_sums_kernel, (const void *)scan_add_kernel, (const void *)dyn_classify_kernel, (const void *)dyn_union_kernel, (const void *)dyn_size_kernel, (const void *)dyn_count_kernel, (const void *)dyn_flag_kernel, (const void *)dyn_scatter_kernel, (const void *)dyn_sortkey_kernel, (const void *)dyn_scatter_vehicle_kernel, (const void *)unpack_pointcloud2_kernel, (const void *)deskew_kernel, (const void *)occ_grid_kernel, (const void *)occ_overlap_kernel};
    for (const void *k : ks) preload_kernel(k);
}


}  // namespace sage
