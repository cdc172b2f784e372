// Shared host/device definitions for the sm_100a SAGE-ICP hot path.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "se3.cuh"

namespace sage {

// ---------------------------------------------------------------------------------------------
// errors / launch accounting
struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct ArgError : std::runtime_error {
    using std::runtime_error::runtime_error;
};
#define SAGE_CUDA(expr)                                                                                      \
    do {                                                                                                     \
        cudaError_t _e = (expr);                                                                             \
        if (_e != cudaSuccess)                                                                               \
            throw ::sage::CudaError(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" __FILE__ ":" + \
                                    std::to_string(__LINE__) + ")");                                         \
    } while (0)

extern std::atomic<long long> g_launches;  // every kernel launch of this library (bench.py: gpu_launches)
#define SAGE_LAUNCH(kernel, grid, block, smem, stream, ...)        \
    do {                                                           \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); \
        ::sage::g_launches.fetch_add(1, std::memory_order_relaxed); \
        SAGE_CUDA(cudaGetLastError());                             \
    } while (0)

// CUDA loads a kernel's code the first time it is launched (lazy module loading, the default since CUDA 12.2).  On a live drive
// that is a stall in the middle of a frame: the first registration whose query count needs another instantiation of the search
// kernel took 150-500 ms on a fresh B200 box (profiles/r02w_small_scans.md §7).  Every translation unit therefore names its kernels
// once, when a map / front end is created on a device: cudaFuncGetAttributes loads the function.
inline void preload_kernel(const void *k) {
    cudaFuncAttributes a;
    if (cudaFuncGetAttributes(&a, k) != cudaSuccess) cudaGetLastError();  // a hint: never an error of the caller
}
void preload_registration_kernels();
void preload_tile_sort_kernels();
void preload_voxel_map_kernels();
void preload_frontend_kernels();

// ---------------------------------------------------------------------------------------------
// growable device / pinned-host buffers
template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    ~DevBuf() {
        if (p) cudaFree(p);
    }
    // grows (content NOT preserved unless keep && stream given)
    void ensure(size_t n, cudaStream_t stream = nullptr, bool keep = false) {
        if (n <= cap) return;
        size_t ncap = cap ? cap : 1;
        while (ncap < n) ncap *= 2;
        T *np = nullptr;
        SAGE_CUDA(cudaMalloc(&np, ncap * sizeof(T)));
        if (keep && p && cap) {
            SAGE_CUDA(cudaMemcpyAsync(np, p, cap * sizeof(T), cudaMemcpyDeviceToDevice, stream));
            SAGE_CUDA(cudaStreamSynchronize(stream));
        }
        if (p) SAGE_CUDA(cudaFree(p));
        p = np, cap = ncap;
    }
};
template <class T>
struct PinBuf {
    T *p = nullptr;
    size_t cap = 0;
    PinBuf() = default;
    PinBuf(const PinBuf &) = delete;
    PinBuf &operator=(const PinBuf &) = delete;
    ~PinBuf() {
        if (p) cudaFreeHost(p);
    }
    void ensure(size_t n) {
        if (n <= cap) return;
        size_t ncap = cap ? cap : 1;
        while (ncap < n) ncap *= 2;
        if (p) SAGE_CUDA(cudaFreeHost(p));
        SAGE_CUDA(cudaMallocHost(&p, ncap * sizeof(T)));
        cap = ncap;
    }
};

// ---------------------------------------------------------------------------------------------
// voxel keys.  Reference: Eigen::Vector3i from static_cast<int>(coord / voxel_size) — truncation toward
// zero of an f64 quotient (core/VoxelHashMap.cpp:52-54,165).  Packed 3 x 21 bits (biased).
constexpr int kKeyBits = 21;
constexpr int kKeyBias = 1 << (kKeyBits - 1);
constexpr uint64_t kEmptyKey = ~0ull;
constexpr uint32_t kNil = 0xffffffffu;

SAGE_HD bool key_in_range(int x, int y, int z) {
    return x > -kKeyBias && x < kKeyBias - 1 && y > -kKeyBias && y < kKeyBias - 1 && z > -kKeyBias && z < kKeyBias - 1;
}
SAGE_HD uint64_t pack_key(int x, int y, int z) {
    return (uint64_t)(uint32_t)(x + kKeyBias) | ((uint64_t)(uint32_t)(y + kKeyBias) << kKeyBits) |
           ((uint64_t)(uint32_t)(z + kKeyBias) << (2 * kKeyBits));
}
SAGE_HD void unpack_key(uint64_t k, int &x, int &y, int &z) {
    const uint64_t m = (1ull << kKeyBits) - 1;
    x = (int)(k & m) - kKeyBias, y = (int)((k >> kKeyBits) & m) - kKeyBias, z = (int)((k >> (2 * kKeyBits)) & m) - kKeyBias;
}
// full-width mixer (the device table is free to use any hash: only set semantics are observable; A.9)
SAGE_HD uint64_t mix64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdull;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ull;
    k ^= k >> 33;
    return k;
}
// static_cast<int>(a / b): true f64 division, truncation toward zero
SAGE_HD int trunc_div(double a, double b) {
#ifdef __CUDA_ARCH__
    return __double2int_rz(__ddiv_rn(a, b));
#else
    return (int)(a / b);
#endif
}

// 16-byte search record of a map point: f32 offsets from the voxel origin (key * voxel_size, so |offset| < voxel_size
// and the rounding error is <= 2^-24 * voxel_size ~ 5e-8 m for 0.8 m voxels) + the label.  The label is stored only when
// it is an integer the f32 holds exactly; anything else becomes NaN, which sends every query that meets the record
// down the exact f64 path.  The search kernel ranks candidates on these records and proves the winner unique within a
// rigorous error band, or re-ranks in f64 on the 32-byte exact records (registration.cu).
#ifdef __CUDACC__
__device__ __forceinline__ float hot_offset(double coord, int key, double voxel_size) {
    return __double2float_rn(__dsub_rn(coord, __dmul_rn((double)key, voxel_size)));
}
__device__ __forceinline__ float hot_label(double label) {
    const float f = __double2float_rn(label);
    return ((double)f == label && truncf(f) == f && fabsf(f) < 8388608.0f) ? f : __int_as_float(0x7fc00000);
}
__device__ __forceinline__ float4 hot_record(const double4 &p, int kx, int ky, int kz, double voxel_size) {
    return make_float4(hot_offset(p.x, kx, voxel_size), hot_offset(p.y, ky, voxel_size), hot_offset(p.z, kz, voxel_size), hot_label(p.w));
}
#endif

// One open-addressing table entry: packed voxel key, block id, and a mirror of the block's point count so a
// probe answers "where and how many" with a single 16-byte load.
struct __align__(16) TblEntry {
    unsigned long long key;
    uint32_t block;
    uint32_t count;
};

// Allocator / statistics block living in device memory (read back asynchronously by the host).
struct MapCtrl {
    uint32_t n_hi;     // blocks ever handed out from the bump pointer
    int32_t n_free;    // entries on the free stack
    uint32_t n_live;   // live voxels
    uint32_t evicted;  // voxels evicted by the last sweep
    uint32_t overflow; // set if a pool/table bound was hit (host guarantees it is not)
    uint32_t range_err; // points dropped because their key was outside the packable range
    unsigned long long n_points;  // scratch for counting kernels
};

// Gauss-Newton state of one registration, device resident so the loop never returns to the host.
struct IcpState {
    Pose est;         // transform the next correspondence kernel applies to `source` (guess, then each estimate)
    Pose T_icp;       // accumulated estimate (core/Registration.cpp:135)
    Pose guess;       // initial guess (for the final T_icp * guess, core/Registration.cpp:140)
    Pose result;      // T_icp * guess once done
    double sums[17];  // 16 normal-equation sums + pair count (as double so one all-reduce carries all)
    double last_norm; // |log(est)| of the last iteration
    double est_th;    // ESTIMATION_THRESHOLD_ (core/Registration.cpp:97)
    int max_iters;    // MAX_NUM_ITERATIONS_ (core/Registration.cpp:96)
    int iter;         // iterations executed
    int done;
    unsigned ticket;  // last-block election
    unsigned comm_error;  // fused peer exchange: a peer did not arrive in time
    unsigned declined;    // tile search: the units are too thinly filled, nothing was done — the host runs the per-query kernel instead
    unsigned long long stat_occupied, stat_candidates;
    unsigned long long stat_scanned, stat_probes, stat_exact, stat_heavy;  // search-kernel work counters (counting launches only)
    unsigned long long stat_staged;  // tile search: records pulled into shared memory by bulk copies
};

// Arguments of the once-per-registration kernels of the tile search (tile_sort.cu) that change from call to call.  They live in
// device memory (one small H2D copy per registration) instead of the kernels' parameter lists, so that the captured CUDA graph of
// those kernels can be replayed unchanged for every scan of the same size.
struct TilePrepArgs {
    Pose guess;           // initial guess (TransformPoints(initial_guess, source), core/Registration.cpp:122)
    const double4 *frame; // the caller's scan
    double est_th;        // ESTIMATION_THRESHOLD_
    int max_iters;        // MAX_NUM_ITERATIONS_
    int apply;            // 1: transform by `guess` while sorting; 0: plain gather (correspondence-only entry points)
};

#ifdef __CUDACC__
// start of a registration (core/Registration.cpp:113-126): estimate = guess, T_icp = identity, counters cleared
__device__ __forceinline__ void icp_state_init(IcpState *st, const Pose &guess, int max_iters, double est_th) {
    st->est = guess;
    st->T_icp = pose_identity();
    st->guess = guess;
    st->result = guess;
    for (int i = 0; i < 17; ++i) st->sums[i] = 0;
    st->last_norm = 0;
    st->est_th = est_th;
    st->max_iters = max_iters;
    st->iter = 0;
    st->done = (max_iters <= 0);
    st->ticket = 0;
    st->stat_occupied = st->stat_candidates = 0;
    st->stat_scanned = st->stat_probes = st->stat_exact = st->stat_heavy = st->stat_staged = 0;
    st->comm_error = 0;
    st->declined = 0;
}
#endif

}  // namespace sage
