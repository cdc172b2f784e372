// Correspondence search + weighted normal equations + on-device Gauss-Newton step (sm_100a).
//
// Replaces, fused into one kernel per iteration:
//   TransformPoints                      core/Registration.cpp:103-111 (applied to `source` at the top of the kernel)
//   VoxelHashMap::GetCorrespondences     core/VoxelHashMap.cpp:48-130
//   AlignClouds (reduce + solve + exp)   core/Registration.cpp:59-94
//   the loop body of RegisterFrame       core/Registration.cpp:127-138
// Correspondences are never materialised: the winner of each query feeds the 16 normal-equation sums directly.
//
// Parallel shape: a group of G lanes (G = 32 by default) owns Q <= G queries per pass; lane j keeps query j's
// transformed point, and the group scans the 27-voxel neighbourhood of one query at a time: lanes probe the
// open-addressed table in parallel (one 16-byte entry load answers block id + count), then read the voxel's
// 32-byte point records with coalesced 256-bit loads, rank them in f64 with exactly the reference's operation
// order, and elect the winner with redux (__reduce_min_sync) on (metric, enumeration order).
#include <cfloat>

#include "nccl_shim.cuh"
#include "voxel_map.cuh"

namespace sage {

constexpr int kNnThreads = 256;
constexpr int kSums = 17;

struct IterParams {
    const TblEntry *tbl;
    uint32_t mask;
    const double4 *blk_pts;
    int stride;
    double voxel_size;
    double4 *src;
    uint32_t n;
    uint32_t chunk;  // Q: queries per group per pass (1..G)
    double max_dist, kern, sem_th;
    IcpState *st;
    double *partials;
    double4 *tgt_out;
    uint8_t *matched_out;
    int apply_est;
    int solve;
    int respect_done;
};

__device__ __forceinline__ double4 ldg256(const double4 *p) {
    double4 v;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ double4 ld256(const double4 *p) {
    double4 v;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(v.x), "=d"(v.y), "=d"(v.z), "=d"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st256(double4 *p, const double4 &v) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}

__device__ __forceinline__ bool tbl_find(const TblEntry *tbl, uint32_t mask, unsigned long long key, uint32_t &block, uint32_t &count) {
    uint32_t i = (uint32_t)mix64(key) & mask;
    while (true) {
        const uint4 e = __ldg(reinterpret_cast<const uint4 *>(tbl + i));
        const unsigned long long k = ((unsigned long long)e.y << 32) | e.x;
        if (k == key) {
            block = e.z, count = e.w;
            return true;
        }
        if (k == kEmptyKey) return false;
        i = (i + 1) & mask;
    }
}

__device__ __forceinline__ double shfl_d(unsigned mask, double v, int lane) { return __shfl_sync(mask, v, lane); }

// One Gauss-Newton step from the reduced sums: x = LDLT(JTJ)^-1 (-JTr); est = exp(x); T_icp = est * T_icp;
// stop when |log(est)| < threshold (core/Registration.cpp:92-93,133-137).
__device__ __noinline__ void icp_solve_step(IcpState *st) {
    const double *S = st->sums;
    // JTJ = sum w [[I, -s^],[s^, |s|^2 I - s s^T]],  JTr = sum w [r; s x r]   
    const double w = S[0], x = S[1], y = S[2], z = S[3], xx = S[4], yy = S[5], zz = S[6], xy = S[7], xz = S[8], yz = S[9];
    double A[6][6] = {{w, 0, 0, 0, z, -y},       {0, w, 0, -z, 0, x},        {0, 0, w, y, -x, 0},
                      {0, -z, y, yy + zz, -xy, -xz}, {z, 0, -x, -xy, xx + zz, -yz}, {-y, x, 0, -xz, -yz, xx + yy}};
    double b[6], xi[6];
    for (int i = 0; i < 6; ++i) b[i] = -S[10 + i];
    solve6_ldlt(A, b, xi);
    const Pose est = pose_exp(xi);
    st->est = est;
    st->T_icp = pose_mul(est, st->T_icp);
    st->iter += 1;
    double lg[6];
    pose_log(est, lg);
    double n2 = 0;
    for (int i = 0; i < 6; ++i) n2 += lg[i] * lg[i];
    const double nrm = sqrt(n2);
    st->last_norm = nrm;
    if (nrm < st->est_th || st->iter >= st->max_iters) {
        st->done = 1;
        st->result = pose_mul(st->T_icp, st->guess);
    }
}

__global__ void icp_init_kernel(IcpState *st, Pose guess, int max_iters, double est_th) {
    st->est = guess;
    st->T_icp = pose_identity();
    st->guess = guess;
    st->result = guess;
    for (int i = 0; i < kSums; ++i) st->sums[i] = 0;
    st->last_norm = 0;
    st->est_th = est_th;
    st->max_iters = max_iters;
    st->iter = 0;
    st->done = (max_iters <= 0);
    st->ticket = 0;
    st->stat_occupied = st->stat_candidates = 0;
}

__global__ void icp_solve_kernel(IcpState *st) {
    if (st->done) return;
    icp_solve_step(st);
}

template <int G, bool STATS>
__global__ void __launch_bounds__(kNnThreads) nn_normal_eq_kernel(IterParams p) {
    __shared__ double s_red[kNnThreads / 32][kSums];
    __shared__ int s_last;
    IcpState *st = p.st;
    if (p.respect_done && st->done) return;

    const Pose est = st->est;
    const int lane = threadIdx.x & 31, gl = lane & (G - 1), gbase = lane - gl;
    const unsigned gmask = (G == 32) ? 0xffffffffu : (((1u << G) - 1u) << gbase);
    const uint32_t group = (blockIdx.x * blockDim.x + threadIdx.x) / G, n_groups = gridDim.x * blockDim.x / G;
    const uint32_t Q = p.chunk;
    const double vs = p.voxel_size, th = p.sem_th;

    double acc[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) acc[k] = 0.0;
    double npairs = 0.0;
    unsigned long long occ = 0, cand = 0;

    for (uint32_t base = group * Q; base < p.n; base += n_groups * Q) {
        const uint32_t q = base + gl;
        const bool valid = gl < Q && q < p.n;
        double sx = 0, sy = 0, sz = 0, sl = 0;
        if (valid) {
            double4 s = ld256(p.src + q);
            if (p.apply_est) {
                pose_act(est, s.x, s.y, s.z, sx, sy, sz);
                sl = s.w;
                st256(p.src + q, make_double4(sx, sy, sz, sl));
            } else {
                sx = s.x, sy = s.y, sz = s.z, sl = s.w;
            }
        }
        const int kx = trunc_div(sx, vs), ky = trunc_div(sy, vs), kz = trunc_div(sz, vs);
        const int ql = __double2int_rz(sl);
        double tx = 0, ty = 0, tz = 0, tl = 0;
        bool ok = false;

        const int nq = min(Q, p.n - base);
        for (int i = 0; i < nq; ++i) {
            const int from = gbase + i;
            const double cx = shfl_d(gmask, sx, from), cy = shfl_d(gmask, sy, from), cz = shfl_d(gmask, sz, from);
            const double cl = shfl_d(gmask, sl, from);
            const int ckx = __shfl_sync(gmask, kx, from), cky = __shfl_sync(gmask, ky, from), ckz = __shfl_sync(gmask, kz, from);
            const int cql = __shfl_sync(gmask, ql, from);

            double best = DBL_MAX;  // closest_distance2 init, core/VoxelHashMap.cpp:81
            uint32_t best_ord = 0xffffffffu, best_idx = 0, ord_base = 0;
#pragma unroll 1
            for (int pr0 = 0; pr0 < 27; pr0 += G) {
                const int pr = pr0 + gl;
                bool found = false;
                uint32_t blk = 0, cnt = 0;
                if (pr < 27) {
                    // enumeration order x outer, y, z inner (core/VoxelHashMap.cpp:57-63)
                    const int nx = ckx + pr / 9 - 1, ny = cky + (pr / 3) % 3 - 1, nz = ckz + pr % 3 - 1;
                    if (key_in_range(nx, ny, nz)) found = tbl_find(p.tbl, p.mask, pack_key(nx, ny, nz), blk, cnt) && cnt > 0;
                }
                unsigned fm = __ballot_sync(gmask, found) & gmask;
                if (STATS && gl == 0) occ += __popc(fm);
                while (fm) {
                    const int l = __ffs(fm) - 1;
                    fm &= fm - 1;
                    const uint32_t b = __shfl_sync(gmask, blk, l), c = __shfl_sync(gmask, cnt, l);
                    const double4 *vp = p.blk_pts + (size_t)b * p.stride;
                    for (uint32_t j = gl; j < c; j += G) {
                        const double4 nb = ldg256(vp + j);
                        const double dx = __dsub_rn(nb.x, cx), dy = __dsub_rn(nb.y, cy), dz = __dsub_rn(nb.z, cz);
                        double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                        // semantic metric, core/VoxelHashMap.cpp:87-88
                        if (__double2int_rz(nb.w) == cql || __double2int_rz(__dmul_rn(nb.w, cl)) == 0) d = __dmul_rn(d, th);
                        if (d < best) best = d, best_ord = ord_base + j, best_idx = b * (uint32_t)p.stride + j;
                    }
                    ord_base += c;
                }
            }
            if (STATS && gl == 0) cand += ord_base;
            // group arg-min on (metric, enumeration order): strict '<' => first minimum wins (core/VoxelHashMap.cpp:89)
            const unsigned long long bits = (unsigned long long)__double_as_longlong(best);
            const uint32_t hi = (uint32_t)(bits >> 32), lo = (uint32_t)bits;
            const uint32_t mhi = __reduce_min_sync(gmask, hi);
            const uint32_t mlo = __reduce_min_sync(gmask, hi == mhi ? lo : 0xffffffffu);
            const bool tie = (hi == mhi) && (lo == mlo);
            const uint32_t mord = __reduce_min_sync(gmask, tie ? best_ord : 0xffffffffu);
            bool accept = false;
            double4 nb = make_double4(0, 0, 0, 0);
            if (mord != 0xffffffffu) {
                const unsigned wm = __ballot_sync(gmask, tie && best_ord == mord) & gmask;
                const uint32_t widx = __shfl_sync(gmask, best_idx, __ffs(wm) - 1);
                nb = ldg256(p.blk_pts + widx);
                const double dx = __dsub_rn(nb.x, cx), dy = __dsub_rn(nb.y, cy), dz = __dsub_rn(nb.z, cz);
                const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                accept = __dsqrt_rn(d2) < p.max_dist;  // core/VoxelHashMap.cpp:111
            }
            if (gl == i) tx = nb.x, ty = nb.y, tz = nb.z, tl = nb.w, ok = accept;
        }

        if (ok) {
            // residual, Geman-McClure weight and the 16 sums (core/Registration.cpp:62-70,79-85)
            const double rx = sx - tx, ry = sy - ty, rz = sz - tz;
            const double r2 = (rx * rx + ry * ry) + rz * rz;
            const double den = p.kern + r2;
            const double w = (p.kern * p.kern) / (den * den);
            const double wx = w * sx, wy = w * sy, wz = w * sz;
            acc[0] += w;
            acc[1] += wx, acc[2] += wy, acc[3] += wz;
            acc[4] += wx * sx, acc[5] += wy * sy, acc[6] += wz * sz;
            acc[7] += wx * sy, acc[8] += wx * sz, acc[9] += wy * sz;
            acc[10] += w * rx, acc[11] += w * ry, acc[12] += w * rz;
            acc[13] += w * (sy * rz - sz * ry), acc[14] += w * (sz * rx - sx * rz), acc[15] += w * (sx * ry - sy * rx);
            npairs += 1.0;
        }
        if (p.tgt_out && valid) {
            st256(p.tgt_out + q, make_double4(tx, ty, tz, tl));
            p.matched_out[q] = ok ? 1 : 0;
        }
    }

    // deterministic reduction: lanes (butterfly) -> warps (fixed order) -> blocks (fixed order, last block)
    const int warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_red[warp][k] = v;
    }
    {
        double v = npairs;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) s_red[warp][16] = v;
    }
    if (STATS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            occ += __shfl_xor_sync(0xffffffffu, occ, o);
            cand += __shfl_xor_sync(0xffffffffu, cand, o);
        }
        if (lane == 0) {
            atomicAdd(&st->stat_occupied, occ);
            atomicAdd(&st->stat_candidates, cand);
        }
    }
    __syncthreads();
    if (threadIdx.x < kSums) {
        double v = 0;
        for (int wv = 0; wv < kNnThreads / 32; ++wv) v += s_red[wv][threadIdx.x];
        p.partials[(size_t)blockIdx.x * kSums + threadIdx.x] = v;
        __threadfence();
    }
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x < kSums) {
        double v = 0;
        const volatile double *part = p.partials;
        for (uint32_t b = 0; b < gridDim.x; ++b) v += part[(size_t)b * kSums + threadIdx.x];
        st->sums[threadIdx.x] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        st->ticket = 0;
        if (p.solve) icp_solve_step(st);
    }
}

// ---------------------------------------------------------------------------------------------
// host side

void VoxelMapGPU::profile_enable(bool on) {
    profile_ = on;
    prof_used_ = 0;
}

void VoxelMapGPU::profile_read(long long *launches, double *ms) {
    set_device();
    SAGE_CUDA(cudaStreamSynchronize(stream_));
    double total = 0;
    for (size_t i = 0; i < prof_used_; ++i) {
        float t = 0;
        SAGE_CUDA(cudaEventElapsedTime(&t, prof_events_[i].first, prof_events_[i].second));
        total += t;
    }
    if (launches) *launches = (long long)prof_used_;
    if (ms) *ms = total;
    prof_used_ = 0;
}

// mode 0: ICP iteration (apply est, solve on device when single rank); mode 1: correspondences/sums of the points
// as given; mode 2: statistics pass.
void VoxelMapGPU::launch_iteration(double4 *src, size_t n, double max_dist, double kernel, double sem_th, int mode,
                                   double4 *tgt_out, uint8_t *matched_out) {
    if (nn_grid_ == 0) {
        int per_sm = 0;
        SAGE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nn_normal_eq_kernel<32, false>, kNnThreads, 0));
        nn_grid_ = sm_count_ * (per_sm > 0 ? per_sm : 1);
        partials_.ensure((size_t)nn_grid_ * kSums);
    }
    constexpr int G = 32;
    const uint32_t groups_max = (uint32_t)nn_grid_ * (kNnThreads / G);
    // queries per group per pass: Q = 1 keeps every SM busy for small scans; for big scans split the queries evenly
    // over the fewest passes so that no group runs an extra, mostly empty pass
    const uint32_t passes = (uint32_t)((n + (size_t)groups_max * G - 1) / ((size_t)groups_max * G));
    uint32_t Q = (uint32_t)((n + (size_t)groups_max * passes - 1) / ((size_t)groups_max * (passes ? passes : 1)));
    Q = Q < 1 ? 1 : (Q > (uint32_t)G ? (uint32_t)G : Q);
    const uint32_t groups_needed = (uint32_t)((n + Q - 1) / Q);
    uint32_t grid = (groups_needed + (kNnThreads / G) - 1) / (kNnThreads / G);
    grid = grid < 1 ? 1 : (grid > (uint32_t)nn_grid_ ? (uint32_t)nn_grid_ : grid);

    IterParams p;
    p.tbl = tbl_.p, p.mask = tbl_cap_ - 1, p.blk_pts = blk_pts_.p, p.stride = stride_, p.voxel_size = voxel_size_;
    p.src = src, p.n = (uint32_t)n, p.chunk = Q;
    p.max_dist = max_dist, p.kern = kernel, p.sem_th = sem_th;
    p.st = icp_.p, p.partials = partials_.p, p.tgt_out = tgt_out, p.matched_out = matched_out;
    p.apply_est = (mode == 0), p.respect_done = (mode == 0);
    p.solve = (mode == 0 && comm_ == nullptr);

    const bool prof = profile_ && mode == 0;
    if (prof) {
        if (prof_used_ == prof_events_.size()) {
            cudaEvent_t a, b;
            SAGE_CUDA(cudaEventCreate(&a));
            SAGE_CUDA(cudaEventCreate(&b));
            prof_events_.emplace_back(a, b);
        }
        SAGE_CUDA(cudaEventRecord(prof_events_[prof_used_].first, stream_));
    }
    if (mode == 2)
        SAGE_LAUNCH((nn_normal_eq_kernel<G, true>), grid, kNnThreads, 0, stream_, p);
    else
        SAGE_LAUNCH((nn_normal_eq_kernel<G, false>), grid, kNnThreads, 0, stream_, p);
    if (prof) {
        SAGE_CUDA(cudaEventRecord(prof_events_[prof_used_].second, stream_));
        ++prof_used_;
    }
    if (mode == 0 && comm_ != nullptr) {
        nccl_allreduce_sum_f64(comm_, icp_.p->sums, kSums, stream_);
        SAGE_LAUNCH(icp_solve_kernel, 1, 1, 0, stream_, icp_.p);
    }
}

int VoxelMapGPU::register_frame_dev(const double4 *frame, size_t n, const Pose &guess, double max_dist, double kernel,
                                    double sem_th, int max_iters, double est_th, Pose &pose_out) {
    set_device();
    if (max_iters <= 0) max_iters = 500;  // MAX_NUM_ITERATIONS_, core/Registration.cpp:96
    if (est_th < 0) est_th = 1e-4;        // ESTIMATION_THRESHOLD_, core/Registration.cpp:97
    // if (voxel_map.Empty()) return initial_guess;  core/Registration.cpp:119
    if (empty()) {
        pose_out = guess;
        return 0;
    }
    icp_.ensure(1);
    icp_pin_.ensure(1);
    src_.ensure(n ? n : 1);
    if (n) SAGE_CUDA(cudaMemcpyAsync(src_.p, frame, n * sizeof(double4), cudaMemcpyDeviceToDevice, stream_));
    SAGE_LAUNCH(icp_init_kernel, 1, 1, 0, stream_, icp_.p, guess, max_iters, est_th);
    int launched = 0;
    while (launched < max_iters) {
        const int batch = (max_iters - launched) < 8 ? (max_iters - launched) : 8;
        for (int b = 0; b < batch; ++b) launch_iteration(src_.p, n, max_dist, kernel, sem_th, 0, nullptr, nullptr);
        launched += batch;
        SAGE_CUDA(cudaMemcpyAsync(icp_pin_.p, icp_.p, sizeof(IcpState), cudaMemcpyDeviceToHost, stream_));
        SAGE_CUDA(cudaStreamSynchronize(stream_));
        if (icp_pin_.p->done) break;
    }
    pose_out = icp_pin_.p->result;
    return icp_pin_.p->iter;
}

double4 *VoxelMapGPU::stage_points(const double *xyzl, size_t n) {
    set_device();
    stage_.ensure(n ? n : 1);
    if (n) SAGE_CUDA(cudaMemcpyAsync(stage_.p, xyzl, n * sizeof(double4), cudaMemcpyHostToDevice, stream_));
    return stage_.p;
}

int VoxelMapGPU::register_frame_host(const double *xyzl, size_t n, const Pose &guess, double max_dist, double kernel,
                                     double sem_th, int max_iters, double est_th, Pose &pose_out) {
    double4 *d = stage_points(xyzl, n);
    return register_frame_dev(d, n, guess, max_dist, kernel, sem_th, max_iters, est_th, pose_out);
}

long long VoxelMapGPU::get_correspondences(const double *xyzl, size_t n, double max_dist, double th, double *target_out,
                                           uint8_t *matched_out) {
    set_device();
    if (n == 0) return 0;
    double4 *d = stage_points(xyzl, n);
    icp_.ensure(1);
    icp_pin_.ensure(1);
    tgt_.ensure(n);
    matched_.ensure(n);
    if (empty()) {
        for (size_t i = 0; i < n; ++i) matched_out[i] = 0;
        return 0;
    }
    SAGE_LAUNCH(icp_init_kernel, 1, 1, 0, stream_, icp_.p, pose_identity(), 1, 0.0);
    launch_iteration(d, n, max_dist, 1.0, th, 1, tgt_.p, matched_.p);
    SAGE_CUDA(cudaMemcpyAsync(target_out, tgt_.p, n * sizeof(double4), cudaMemcpyDeviceToHost, stream_));
    SAGE_CUDA(cudaMemcpyAsync(matched_out, matched_.p, n, cudaMemcpyDeviceToHost, stream_));
    SAGE_CUDA(cudaMemcpyAsync(icp_pin_.p, icp_.p, sizeof(IcpState), cudaMemcpyDeviceToHost, stream_));
    SAGE_CUDA(cudaStreamSynchronize(stream_));
    return (long long)icp_pin_.p->sums[16];
}

void VoxelMapGPU::normal_equations(const double *xyzl, size_t n, double max_dist, double kernel, double sem_th, double JTJ[36],
                                   double JTr[6], long long *pairs) {
    set_device();
    double4 *d = stage_points(xyzl, n);
    icp_.ensure(1);
    icp_pin_.ensure(1);
    SAGE_LAUNCH(icp_init_kernel, 1, 1, 0, stream_, icp_.p, pose_identity(), 1, 0.0);
    if (!empty() && n) launch_iteration(d, n, max_dist, kernel, sem_th, 1, nullptr, nullptr);
    SAGE_CUDA(cudaMemcpyAsync(icp_pin_.p, icp_.p, sizeof(IcpState), cudaMemcpyDeviceToHost, stream_));
    SAGE_CUDA(cudaStreamSynchronize(stream_));
    const double *S = icp_pin_.p->sums;
    const double w = S[0], x = S[1], y = S[2], z = S[3], xx = S[4], yy = S[5], zz = S[6], xy = S[7], xz = S[8], yz = S[9];
    const double A[6][6] = {{w, 0, 0, 0, z, -y},       {0, w, 0, -z, 0, x},        {0, 0, w, y, -x, 0},
                            {0, -z, y, yy + zz, -xy, -xz}, {z, 0, -x, -xy, xx + zz, -yz}, {-y, x, 0, -xz, -yz, xx + yy}};
    for (int i = 0; i < 6; ++i) {
        for (int j = 0; j < 6; ++j) JTJ[i * 6 + j] = A[i][j];
        JTr[i] = S[10 + i];
    }
    if (pairs) *pairs = (long long)S[16];
}

void VoxelMapGPU::nn_stats(const double *xyzl, size_t n, unsigned long long *occupied, unsigned long long *candidates) {
    set_device();
    *occupied = *candidates = 0;
    if (n == 0 || empty()) return;
    double4 *d = stage_points(xyzl, n);
    icp_.ensure(1);
    icp_pin_.ensure(1);
    SAGE_LAUNCH(icp_init_kernel, 1, 1, 0, stream_, icp_.p, pose_identity(), 1, 0.0);
    launch_iteration(d, n, 0.0, 1.0, 1.0, 2, nullptr, nullptr);
    SAGE_CUDA(cudaMemcpyAsync(icp_pin_.p, icp_.p, sizeof(IcpState), cudaMemcpyDeviceToHost, stream_));
    SAGE_CUDA(cudaStreamSynchronize(stream_));
    *occupied = icp_pin_.p->stat_occupied;
    *candidates = icp_pin_.p->stat_candidates;
}

}  // namespace sage
