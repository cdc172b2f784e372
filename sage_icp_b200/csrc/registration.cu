// Correspondence search + weighted normal equations + on-device Gauss-Newton step (sm_100a).
//
// Replaces, fused into one kernel per iteration:
//   TransformPoints                      core/Registration.cpp:103-111 (applied to `source` at the top of the kernel)
//   VoxelHashMap::GetCorrespondences     core/VoxelHashMap.cpp:48-130
//   AlignClouds (reduce + solve + exp)   core/Registration.cpp:59-94
//   the loop body of RegisterFrame       core/Registration.cpp:127-138
// Correspondences are never materialised: the winner of each query feeds the 16 normal-equation sums directly.
//
// Search strategy (results are IDENTICAL to the reference's f64 scan of all 27 voxels):
//   1. A group of G lanes owns G queries per pass; lane j keeps query j (f64, transformed in place exactly as the
//      reference does) and the group searches one query at a time.
//   2. Candidates are ranked on 16-byte f32 voxel-relative records (common.cuh: hot_record) — half the bytes of the
//      reference's Vector4d and f32 arithmetic.  Every f32 metric carries a rigorous error bound e(D).
//   3. Home voxel first; a neighbour voxel is probed and scanned only if a conservative lower bound of the metric over
//      its bounding box does not already exceed the best metric found (exact: such a voxel cannot hold the arg-min,
//      nor tie with it).
//   4. The f32 arg-min is accepted only if no other candidate lies within the error band; otherwise (rare: ~1e-4 of
//      the queries) the query is re-ranked by nn_exact(): the reference's own f64 operation sequence over all 27 voxels in
//      enumeration order with the strict-'<' first-wins rule (core/VoxelHashMap.cpp:57-63,81-93).
//   5. The winner's exact f64 record is fetched for the acceptance test and the residual.
#include <cooperative_groups.h>

#include <cfloat>
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "nccl_shim.cuh"
#include "voxel_map.cuh"

namespace sage {

#ifndef SAGE_NN_THREADS
#define SAGE_NN_THREADS 256
#endif
constexpr int kNnThreads = SAGE_NN_THREADS;
static_assert(kNnThreads % 64 == 0 && kNnThreads >= 64, "whole warps, and icp_step_block needs two of them");
constexpr int kSums = 17;
constexpr int kDbg = 12;  // debug timeline stamps per block
constexpr int kMaxPeers = 8;   // GPUs of one NVSwitch box
constexpr int kXchgSlot = 24;  // doubles per (parity, source rank) slot: 17 sums + tag, padded to 192 bytes


struct IterParams {
    const TblEntry *tbl;
    uint32_t mask;
    const double4 *blk_pts;
    const float4 *blk_hot;
    int stride;
    double voxel_size;
    double4 *src;
    uint32_t n;
    double max_dist, kern, sem_th;
    // f32 search constants (host-computed): voxel size, metric scale for compatible labels, error model
    // e(D) = err_a * sqrt(D) + err_b * D + err_c on a metric whose squared distance is D, lower-bound shrink + margin
    float vs32, th32, smin32, inv_smin32, err_scale, err_a, err_b, err_c, box_margin;
    int fast_ok;  // 0: sem_th outside (0, inf) or non-finite -> every query takes the exact path
    IcpState *st;
    double *partials;
    double4 *tgt_out;
    uint8_t *matched_out;
    int apply_est;
    int solve;
    int respect_done;
    int all_warp;      // 1: warp-per-query for every query (small scans)
    // fused all-reduce over NVLink peer memory (multi-GPU): xchg_peer[k] = rank k's exchange buffer mapped here (CUDA IPC),
    // laid out [parity][source rank][kXchgSlot]; xchg_tag = this launch's sequence number (same on every rank)
    int xchg_world, xchg_rank;
    double *xchg_peer[kMaxPeers];
    unsigned long long xchg_tag;
    unsigned long long xchg_timeout_ns;  // how long the last block waits for the slowest rank before it gives up (comm_error)
    int light_probes;  // neighbour probes a query may spend in the thread-per-query phase before it is deferred
    int step_everywhere;  // persistent kernels, single rank: every block reduces the partials and takes the step (LoopState)
    // tile search (search_tile.cuh): unit boundaries in the sorted query array [n_units + 1], their number (device scalar), and
    // the capacity of the block's staging area in 16-byte records
    const uint32_t *tile_units, *tile_n_units, *tile_perm;  // tile_perm: sorted position -> index in the caller's array
    const uint32_t *tile_order;  // hand-out order of the units (largest first); null = list order
    double *tile_unit_part;    // [unit][17]
    uint32_t tile_fill;        // mean queries per unit below which the kernel declines (0: never)
    uint32_t *tile_ctl;        // [0] units handed out so far in this registration, [32] its value at the start of the iteration
    uint32_t *tile_group_cnt;  // [group]
    uint32_t tile_stage_cap;
    unsigned long long *dbg;  // optional per-block timeline (tools/perf_probe.py): 4 globaltimer stamps per block + 4 global
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

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ double shfl_d(unsigned mask, double v, int lane) { return __shfl_sync(mask, v, lane); }

// One Gauss-Newton step from the reduced sums: x = LDLT(JTJ)^-1 (-JTr); est = exp(x); T_icp = est * T_icp; stop when
// |log(est)| < threshold (core/Registration.cpp:92-93,133-137).  icp_step_block is called by EVERY thread of a block of >= 64
// threads (it synchronises).
__device__ __noinline__ void icp_solve_xi(const double *S, double xi[6]) {
    // Normal equations with s = sum w s, r_t = sum w r, r_r = sum w (s x r):
    //   JTJ = [[ w I, -[s]x ], [ [s]x, C ]],  C = sum w (|s|^2 I - s s^T),   JTJ (u, o) = -(r_t, r_r).
    // The translation block is w I, so the 6x6 system collapses exactly onto its 3x3 Schur complement
    //   (C - (|s|^2 I - s s^T) / w) o = -r_r + (s x r_t) / w,      u = (-r_t - o x s) / w,
    // i.e. the second moments about the weighted centroid: better conditioned than the raw 6x6 (whose entries grow like
    // |s|^2) and ~5 us cheaper than a pivoted 6x6 LDLT in one thread.  Same solution as the reference's
    // JTJ.ldlt().solve(-JTr) (core/Registration.cpp:92) up to rounding (~1e-12 relative; tolerance is 1e-4 m / 1e-5 rad).
    const double w = S[0], x = S[1], y = S[2], z = S[3], xx = S[4], yy = S[5], zz = S[6], xy = S[7], xz = S[8], yz = S[9];
    const double btx = -S[10], bty = -S[11], btz = -S[12], brx = -S[13], bry = -S[14], brz = -S[15];
#pragma unroll
    for (int i = 0; i < 6; ++i) xi[i] = 0.0;
    if (!(w > 0.0)) return;  // no correspondence: zero step, as the reference's LDLT of a zero matrix
    const double iw = 1.0 / w;
    const double m00 = (yy + zz) - (y * y + z * z) * iw, m11 = (xx + zz) - (x * x + z * z) * iw, m22 = (xx + yy) - (x * x + y * y) * iw;
    const double m01 = -xy + x * y * iw, m02 = -xz + x * z * iw, m12 = -yz + y * z * iw;
    // rhs = b_r - (s x b_t) / w
    const double qx = brx - (y * btz - z * bty) * iw, qy = bry - (z * btx - x * btz) * iw, qz = brz - (x * bty - y * btx) * iw;
    // symmetric 3x3 solve by cofactors
    const double c00 = m11 * m22 - m12 * m12, c01 = m02 * m12 - m01 * m22, c02 = m01 * m12 - m02 * m11;
    const double c11 = m00 * m22 - m02 * m02, c12 = m01 * m02 - m00 * m12, c22 = m00 * m11 - m01 * m01;
    const double det = m00 * c00 + m01 * c01 + m02 * c02;
    // Rank-deficient rotation block (one or two correspondences, collinear points: track loss, a scan leaving the map): det is
    // rounding noise there, not 0, and dividing by it would throw the pose far away.  The reference's LDLT divides by whatever
    // pivot rounding leaves (Eigen zeroes only pivots below 1/highest()), so its step is noise-dominated too; here the step
    // degrades to the translation that aligns the weighted centroids (o = 0, u = b_t / w), which is bounded by the residuals.
    // Two ways of being singular up to rounding: the centred second moments M are entirely cancellation noise of the uncentred
    // ones (all sources coincide: tr M <= 1e-13 * sum w |s|^2), or M itself has a null direction (sources on a line: |det M| <=
    // 1e-12 (tr M / 3)^3).  A well-conditioned system is twelve orders of magnitude away from either.
    const double tr = (m00 + m11 + m22) * (1.0 / 3.0);
    if (!(tr > 1e-13 * ((xx + yy) + zz)) || !(fabs(det) > 1e-12 * tr * tr * tr)) {
        xi[0] = btx * iw, xi[1] = bty * iw, xi[2] = btz * iw;
        return;
    }
    const double id = 1.0 / det;
    const double ox = (c00 * qx + c01 * qy + c02 * qz) * id, oy = (c01 * qx + c11 * qy + c12 * qz) * id, oz = (c02 * qx + c12 * qy + c22 * qz) * id;
    // u = (b_t - o x s) / w
    xi[0] = (btx - (oy * z - oz * y)) * iw, xi[1] = (bty - (oz * x - ox * z)) * iw, xi[2] = (btz - (ox * y - oy * x)) * iw;
    xi[3] = ox, xi[4] = oy, xi[5] = oz;
}
// IcpState fields that change from one iteration to the next are read past L1 (ld.global.cg): in the persistent kernel the
// block that takes the step is a different one every iteration, on an SM whose L1 may still hold the line from an earlier turn.
__device__ __forceinline__ Pose load_pose_cg(const Pose *q) {
    const double *d = reinterpret_cast<const double *>(q);
    return Pose{__ldcg(d), __ldcg(d + 1), __ldcg(d + 2), __ldcg(d + 3), __ldcg(d + 4), __ldcg(d + 5), __ldcg(d + 6)};
}
// The step is a chain of dependent f64 operations (two divisions, sqrt, sin/cos, atan2: ~8 us in one thread), so the independent
// pieces run side by side in two warps: [solve] -> [rotation part of exp | translation part of exp] -> [pose product | log norm]
// -> [stop test; the final pose only when the loop ends].  Same formulas, same operation order as pose_exp / pose_log.
// Pure part of the step: sums -> estimate (*s_pose) and |log(estimate)| (*s_norm), nothing but shared memory touched.
__device__ __forceinline__ void icp_step_local(const double *sums, Pose *s_pose, double *s_norm, double est_th, unsigned long long *dbg = nullptr) {
    __shared__ double s_xi[6];
    __shared__ int s_need_log;
    if (threadIdx.x == 0) {
        double xi[6];
        icp_solve_xi(sums, xi);
#pragma unroll
        for (int i = 0; i < 6; ++i) s_xi[i] = xi[i];
        // The stop test needs |log(exp(xi))| (core/Registration.cpp:137).  For a rotation below 1 rad that IS |xi| up to the rounding
        // of the exp / log round trip (< 1e-13 relative), and pose_log — atan2, a square root, three divisions in one dependent f64
        // chain — was the longest piece of the step (3.3-4.4 us of 10, profiles/r02_tile_kernel.md).  So the test is made on |xi|
        // whenever |xi| is further than 1e-9 (relative) from the threshold, where both norms fall on the same side of it; only a step
        // that lands within that sliver of the threshold (or a non-finite / large-rotation one) still takes the full logarithm below.
        const double n = sqrt((xi[0] * xi[0] + xi[1] * xi[1] + xi[2] * xi[2]) + (xi[3] * xi[3] + xi[4] * xi[4] + xi[5] * xi[5]));
        const double th2 = (xi[3] * xi[3] + xi[4] * xi[4]) + xi[5] * xi[5];
        *s_norm = n;
        s_need_log = (th2 < 1.0 && fabs(n - est_th) > 1e-9 * fmax(n, est_th)) ? 0 : 1;
        if (dbg) dbg[5] = gtime();
    }
    __syncthreads();
    if (threadIdx.x == 0 || threadIdx.x == 32) {  // Sophus SE3::exp, tangent = (upsilon, omega) — se3.cuh pose_exp, split in two
        const double eps = 1e-10;
        const double wx = s_xi[3], wy = s_xi[4], wz = s_xi[5];
        const double th2 = (wx * wx + wy * wy) + wz * wz;
        if (threadIdx.x == 0) {
            double imag, real;
            if (th2 < eps * eps) {
                const double th4 = th2 * th2;
                imag = 0.5 - (1.0 / 48.0) * th2 + (1.0 / 3840.0) * th4;
                real = 1.0 - (1.0 / 8.0) * th2 + (1.0 / 384.0) * th4;
            } else {
                const double theta = sqrt(th2), half = 0.5 * theta;
                imag = sin(half) / theta;
                real = cos(half);
            }
            double qw = real, qx = imag * wx, qy = imag * wy, qz = imag * wz;
            const double n = sqrt(qw * qw + qx * qx + qy * qy + qz * qz);
            s_pose->qw = qw / n, s_pose->qx = qx / n, s_pose->qy = qy / n, s_pose->qz = qz / n;
        } else {
            const double theta = th2 < eps * eps ? 0.0 : sqrt(th2);
            double a, b;
            if (theta < eps) {
                a = 0.5, b = 1.0 / 6.0;
            } else {
                a = (1.0 - cos(theta)) / th2;
                b = (theta - sin(theta)) / (th2 * theta);
            }
            const double ux = s_xi[0], uy = s_xi[1], uz = s_xi[2];
            const double c1x = wy * uz - wz * uy, c1y = wz * ux - wx * uz, c1z = wx * uy - wy * ux;
            const double c2x = wy * c1z - wz * c1y, c2y = wz * c1x - wx * c1z, c2z = wx * c1y - wy * c1x;
            s_pose->tx = ux + a * c1x + b * c2x;
            s_pose->ty = uy + a * c1y + b * c2y;
            s_pose->tz = uz + a * c1z + b * c2z;
        }
    }
    __syncthreads();
    if (dbg && threadIdx.x == 0) dbg[6] = gtime();
    if (s_need_log && threadIdx.x == 32) {
        double lg[6];
        pose_log(*s_pose, lg);
        double n2 = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) n2 += lg[i] * lg[i];
        *s_norm = sqrt(n2);
    }
}
// One step on the device-resident state, by the block that finishes an iteration.  What thread 0 needs from global memory for the
// bookkeeping — the accumulated estimate, the iteration count, the limits — is requested before the solve, so that the ~1 us of each
// of those loads passes during it instead of after it (they were a chain of three round trips behind the step: 3.3 us of 7.4).
__device__ __forceinline__ void icp_step_block(IcpState *st, const double *sums, Pose *s_pose, double *s_norm, unsigned long long *dbg = nullptr) {
    Pose T_prev = pose_identity();
    int it_prev = 0, max_it = 0;
    if (threadIdx.x == 0) {
        T_prev = load_pose_cg(&st->T_icp);
        it_prev = __ldcg(&st->iter);
        max_it = st->max_iters;
    }
    const double est_th = st->est_th;  // constant during a registration
    icp_step_local(sums, s_pose, s_norm, est_th, dbg);
    if (threadIdx.x == 0) {  // beside the (rare) log of icp_step_local in thread 32
        st->est = *s_pose;
        T_prev = pose_mul(*s_pose, T_prev);  // T_icp = est * T_icp, core/Registration.cpp:135
        st->T_icp = T_prev;
        st->iter = it_prev + 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        st->last_norm = *s_norm;
        if (*s_norm < est_th || it_prev + 1 >= max_it) {
            st->result = pose_mul(T_prev, st->guess);  // T_icp * guess (core/Registration.cpp:140): needed once, when the loop ends
            st->done = 1;
        }
    }
}

__global__ void icp_init_kernel(IcpState *st, Pose guess, int max_iters, double est_th) { icp_state_init(st, guess, max_iters, est_th); }

__global__ void icp_solve_kernel(IcpState *st) {  // <<<1, 64>>>, after the NCCL all-reduce of the sums
    __shared__ Pose s_pose;
    __shared__ double s_norm;
    if (st->done) return;
    icp_step_block(st, st->sums, &s_pose, &s_norm);
}

// ---------------------------------------------------------------------------------------------
// Exact path: the reference's f64 scan of all 27 voxels in enumeration order (x outer, y, z inner; stored order inside
// a voxel), strict '<' from DBL_MAX.  Group-cooperative; returns the winner's record index (block * stride + slot) or
// kNil when the neighbourhood holds no point, identical on every lane of the group.
template <int G>
__device__ __noinline__ uint32_t nn_exact(const IterParams &p, unsigned gmask, int gl, double cx, double cy, double cz, double cl,
                                          int ckx, int cky, int ckz) {
    const int cql = __double2int_rz(cl);
    const double th = p.sem_th;
    double best = DBL_MAX;  // closest_distance2 init, core/VoxelHashMap.cpp:81
    uint32_t best_ord = 0xffffffffu, best_idx = kNil;
#pragma unroll 1
    for (int pr0 = 0; pr0 < 27; pr0 += G) {
        const int pr = pr0 + gl;
        bool found = false;
        uint32_t blk = 0, cnt = 0;
        if (pr < 27) {
            const int nx = ckx + pr / 9 - 1, ny = cky + (pr / 3) % 3 - 1, nz = ckz + pr % 3 - 1;
            if (key_in_range(nx, ny, nz)) found = tbl_find(p.tbl, p.mask, pack_key(nx, ny, nz), blk, cnt) && cnt > 0;
        }
        unsigned fm = __ballot_sync(gmask, found) & gmask;
        // up to four found voxels per round: their 32-byte loads are independent, so a round costs one trip to memory instead of
        // four.  Candidates therefore arrive out of enumeration order; the running minimum is kept lexicographically on (metric,
        // enumeration index: voxel, slot), which is exactly what the reference's strict '<' over the enumeration yields.
        while (fm) {
            uint32_t b[4], c[4], ob[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int l = fm ? __ffs(fm) - 1 : -1;
                fm &= fm - 1;  // 0 & anything stays 0
                const int src = l < 0 ? 0 : l;
                b[u] = __shfl_sync(gmask, blk, src);
                c[u] = l < 0 ? 0u : __shfl_sync(gmask, cnt, src);
                ob[u] = (uint32_t)(pr0 + (src & (G - 1))) << 16;  // voxel enumeration index : slot
            }
            const uint32_t cmax = max(max(c[0], c[1]), max(c[2], c[3]));
            for (uint32_t j = gl; j < cmax; j += G) {
                double4 nb[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (j < c[u]) nb[u] = ldg256(p.blk_pts + (size_t)b[u] * p.stride + j);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (j < c[u]) {
                        const double dx = __dsub_rn(nb[u].x, cx), dy = __dsub_rn(nb[u].y, cy), dz = __dsub_rn(nb[u].z, cz);
                        double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                        // semantic metric, core/VoxelHashMap.cpp:87-88
                        if (__double2int_rz(nb[u].w) == cql || __double2int_rz(__dmul_rn(nb[u].w, cl)) == 0) d = __dmul_rn(d, th);
                        const uint32_t ord = ob[u] + j;
                        // (best_ord != ~0: DBL_MAX itself never wins in the reference either — its scan starts from DBL_MAX with '<')
                        if (d < best || (d == best && best_ord != 0xffffffffu && ord < best_ord))
                            best = d, best_ord = ord, best_idx = b[u] * (uint32_t)p.stride + j;
                    }
            }
        }
    }
    // group arg-min on (metric, enumeration order): strict '<' => first minimum wins (core/VoxelHashMap.cpp:89).
    // Plain shuffle butterfly on f64 (this path is rare, and it stays correct for a negative sem_th).
#pragma unroll
    for (int s = G / 2; s > 0; s >>= 1) {
        const double om = __shfl_xor_sync(gmask, best, s);
        const uint32_t oo = __shfl_xor_sync(gmask, best_ord, s), oi = __shfl_xor_sync(gmask, best_idx, s);
        if (oo != 0xffffffffu && (best_ord == 0xffffffffu || om < best || (om == best && oo < best_ord))) best = om, best_ord = oo, best_idx = oi;
    }
    return best_ord == 0xffffffffu ? kNil : best_idx;
}

// ---------------------------------------------------------------------------------------------
// f32 ranking helpers

// e(D): bound on |m32 - m| for a metric whose squared distance is <= D (derivation in DESIGN.md §4)
__device__ __forceinline__ float err_of(const IterParams &p, float D) { return p.err_scale * (p.err_a * sqrtf(D) + p.err_b * D) + p.err_c; }
// A voxel whose metric lower bound exceeds this cannot hold the f64 arg-min or tie it: exact best-so-far <= gb + e(gb/smin)
__device__ __forceinline__ float prune_bound(const IterParams &p, float gb) { return gb + 2.0f * err_of(p, gb * p.inv_smin32); }
// The f64 arg-min j* satisfies m32(j*) <= gm + e(D_min) + e(D_j*) with D_min <= gm/smin, D_j* <= (gm + e(D_min))/smin
__device__ __forceinline__ float band_limit(const IterParams &p, float gm) {
    const float e1 = err_of(p, gm * p.inv_smin32);
    return gm + 1.01f * (e1 + err_of(p, (gm + e1) * p.inv_smin32));
}

// rank one search record against the query at (rx, ry, rz) relative to the record's voxel origin
__device__ __forceinline__ void rank_record(const float4 h, uint32_t idx, float rx, float ry, float rz, float qlf, float th32, float &min1,
                                            float &min2, uint32_t &idx1, bool &odd) {
    const float dx = h.x - rx, dy = h.y - ry, dz = h.z - rz;
    const float D = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    // labels are exact integers here (hot_label): int(l_n) == int(l_q) || int(l_n * l_q) == 0  <=>  equal or one is 0
    const bool compat = (h.w == qlf) || (h.w * qlf == 0.0f);
    odd |= (h.w != h.w);  // non-integer label stored as NaN: the f32 ranking is not valid for this query
    const float m = compat ? D * th32 : D;
    if (m < min1) {
        min2 = min1, min1 = m, idx1 = idx;
    } else {
        min2 = fminf(min2, m);
    }
}

// one thread scans one voxel (8 independent 16-byte loads in flight)
#ifndef SAGE_SCAN_UNROLL
#define SAGE_SCAN_UNROLL 4
#endif
#ifndef SAGE_LIGHT_MINB
#define SAGE_LIGHT_MINB 4
#endif
// one thread scans one voxel, SAGE_SCAN_UNROLL independent 16-byte loads in flight.  (Measured on B200: deeper unrolling
// or a software pipeline across batches costs more in spills than it hides — profiles/r01b_search_kernel_timeline.md.)
__device__ __forceinline__ void scan_voxel_thread(const float4 *__restrict__ hot, uint32_t base_idx, uint32_t cnt, float rx, float ry,
                                                  float rz, float qlf, float th32, float &min1, float &min2, uint32_t &idx1, bool &odd) {
    constexpr int U = SAGE_SCAN_UNROLL;
    uint32_t j = 0;
    for (; j + U <= cnt; j += U) {
        float4 h[U];
#pragma unroll
        for (int u = 0; u < U; ++u) h[u] = __ldg(hot + base_idx + j + u);
#pragma unroll
        for (int u = 0; u < U; ++u) rank_record(h[u], base_idx + j + u, rx, ry, rz, qlf, th32, min1, min2, idx1, odd);
    }
    if (j < cnt) {
        float4 h[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (j + u < cnt) h[u] = __ldg(hot + base_idx + j + u);
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (j + u < cnt) rank_record(h[u], base_idx + j + u, rx, ry, rz, qlf, th32, min1, min2, idx1, odd);
    }
}

// Per-axis squared, scaled distance from the query (offset r from its home voxel's origin, home key k) to the boxes of
// the two face neighbours k-1 / k+1.  A voxel with key k holds coordinates in [0, vs) (k > 0), (-vs, vs) (k = 0) or
// (-vs, 0] (k < 0) relative to k * vs (truncation toward zero); `margin` absorbs every rounding.
__device__ __forceinline__ void axis_bounds(float r, int k, float vs, float margin, float smin, float &sm, float &sp) {
    const float em = fmaxf((r + vs) - (k - 1 >= 0 ? vs : 0.0f) - margin, 0.0f);
    const float ep = fmaxf((k + 1 <= 0 ? -vs : 0.0f) - (r - vs) - margin, 0.0f);
    sm = smin * em * em, sp = smin * ep * ep;
}

// residual, Geman-McClure weight and the 16 sums + pair count of one accepted pair, into column t of the shared sums
// (core/Registration.cpp:62-70,79-85)
template <int COLS>
__device__ __forceinline__ void accumulate_pair(double (*s_acc)[COLS], int t, double kern, double sx, double sy, double sz, double tx,
                                                double ty, double tz) {
    const double rx = sx - tx, ry = sy - ty, rz = sz - tz;
    const double r2 = (rx * rx + ry * ry) + rz * rz;
    const double den = kern + r2;
    const double w = (kern * kern) / (den * den);
    const double wx = w * sx, wy = w * sy, wz = w * sz;
    s_acc[0][t] += w;
    s_acc[1][t] += wx, s_acc[2][t] += wy, s_acc[3][t] += wz;
    s_acc[4][t] += wx * sx, s_acc[5][t] += wy * sy, s_acc[6][t] += wz * sz;
    s_acc[7][t] += wx * sy, s_acc[8][t] += wx * sz, s_acc[9][t] += wy * sz;
    s_acc[10][t] += w * rx, s_acc[11][t] += w * ry, s_acc[12][t] += w * rz;
    s_acc[13][t] += w * (sy * rz - sz * ry), s_acc[14][t] += w * (sz * rx - sx * rz), s_acc[15][t] += w * (sx * ry - sy * rx);
    s_acc[16][t] += 1.0;
}

// acceptance test on the exact records: ||n - p|| < max_correspondance_distance (core/VoxelHashMap.cpp:111)
__device__ __forceinline__ bool accept_pair(const double4 &nb, double sx, double sy, double sz, double max_dist) {
    const double dx = __dsub_rn(nb.x, sx), dy = __dsub_rn(nb.y, sy), dz = __dsub_rn(nb.z, sz);
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
    return __dsqrt_rn(d2) < max_dist;
}

// ---------------------------------------------------------------------------------------------
// Warp-cooperative f32 search of one query over all 27 voxels: lane = voxel in the reference's enumeration order (all 27
// probed at once), found voxels scanned four at a time by the whole warp, same ranking + error band; ambiguous queries
// fall through to nn_exact<32>.  Returns the winner's record index or kNil, identical on every lane.
// What the thread-per-query phase hands over with a deferred query: the best two metrics and the winner so far, and which of
// the 27 voxels it has already scanned.  {INF, INF, kNil, 0} = nothing done yet (small-scan mode).
struct Carried {
    float min1, min2;
    uint32_t idx1, visited;
};

template <bool COUNT>
__device__ __noinline__ uint32_t search_query_warp(const IterParams &p, int lane, const double4 s, const Carried cs, unsigned long long &n_scanned,
                                                   unsigned long long &n_probes, unsigned long long &n_exact) {
    const double vs = p.voxel_size;
    const float vs32 = p.vs32, th32 = p.th32;
    const float INF = __int_as_float(0x7f800000);
    const unsigned FULL = 0xffffffffu;
    const int kx = trunc_div(s.x, vs), ky = trunc_div(s.y, vs), kz = trunc_div(s.z, vs);
    const float bx = hot_offset(s.x, kx, vs), by = hot_offset(s.y, ky, vs), bz = hot_offset(s.z, kz, vs);
    const float qlf = hot_label(s.w);
    bool odd = !p.fast_ok || (qlf != qlf) || !(fabsf(bx) <= 2.0f * vs32 && fabsf(by) <= 2.0f * vs32 && fabsf(bz) <= 2.0f * vs32);
    uint32_t widx = kNil;
    if (!odd) {
        // lane = voxel in the reference's enumeration order.  Skip what the first phase scanned and what its best-so-far
        // already rules out (same bound as there: lower bound of the metric over the voxel's box > best + 2 e).
        const int ox = lane / 9 - 1, oy = (lane / 3) % 3 - 1, oz = lane % 3 - 1;
        bool found = false;
        uint32_t blk = 0, cnt = 0;
        if (lane < 27 && !((cs.visited >> lane) & 1u)) {
            float sxm, sxp, sym, syp, szm, szp;
            axis_bounds(bx, kx, vs32, p.box_margin, p.smin32, sxm, sxp);
            axis_bounds(by, ky, vs32, p.box_margin, p.smin32, sym, syp);
            axis_bounds(bz, kz, vs32, p.box_margin, p.smin32, szm, szp);
            const float lb = (ox < 0 ? sxm : (ox > 0 ? sxp : 0.0f)) + (oy < 0 ? sym : (oy > 0 ? syp : 0.0f)) + (oz < 0 ? szm : (oz > 0 ? szp : 0.0f));
            const int nx = kx + ox, ny = ky + oy, nz = kz + oz;
            if (lb <= prune_bound(p, cs.min1) && key_in_range(nx, ny, nz)) {
                found = tbl_find(p.tbl, p.mask, pack_key(nx, ny, nz), blk, cnt) && cnt > 0;
                if (COUNT) n_probes += 1;
            }
        }
        unsigned fm = __ballot_sync(FULL, found);
        // lane 0 carries the first phase's result into the reduction
        float min1 = lane == 0 ? cs.min1 : INF, min2 = lane == 0 ? cs.min2 : INF;
        uint32_t idx1 = lane == 0 ? cs.idx1 : kNil;
        while (fm) {  // four voxels per round: their loads are independent
            int l[4];
            uint32_t b[4], c[4];
            float x[4], y[4], z[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                l[u] = fm ? __ffs(fm) - 1 : -1;
                fm &= fm - 1;  // 0 & anything stays 0
                const int src = l[u] < 0 ? 0 : l[u];
                b[u] = __shfl_sync(FULL, blk, src) * (uint32_t)p.stride;
                c[u] = l[u] < 0 ? 0u : __shfl_sync(FULL, cnt, src);
                x[u] = bx - (float)(src / 9 - 1) * vs32, y[u] = by - (float)((src / 3) % 3 - 1) * vs32, z[u] = bz - (float)(src % 3 - 1) * vs32;
            }
            const uint32_t cmax = max(max(c[0], c[1]), max(c[2], c[3]));
            for (uint32_t j = lane; j < cmax; j += 32) {
                float4 h[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (j < c[u]) h[u] = __ldg(p.blk_hot + b[u] + j);
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (j < c[u]) rank_record(h[u], b[u] + j, x[u], y[u], z[u], qlf, th32, min1, min2, idx1, odd);
            }
            if (COUNT && lane == 0) n_scanned += c[0] + c[1] + c[2] + c[3];
        }
        odd = __any_sync(FULL, odd);
        if (!odd) {
            const float gm = __uint_as_float(__reduce_min_sync(FULL, __float_as_uint(min1)));
            if (gm < INF) {
                const float T = band_limit(p, gm);
                const unsigned in1 = __ballot_sync(FULL, min1 <= T), in2 = __ballot_sync(FULL, min2 <= T);
                if (__popc(in1) == 1 && in2 == 0)
                    widx = __shfl_sync(FULL, idx1, __ffs(in1) - 1);
                else
                    odd = true;
            }
        }
    }
    if (odd) {
        widx = nn_exact<32>(p, FULL, lane, s.x, s.y, s.z, s.w, kx, ky, kz);
        if (COUNT && lane == 0) n_exact += 1;
    }
    return widx;
}

// Acceptance, residual and sums of up to 32 warp-searched queries at once: lane k holds the k-th query of the batch.
template <int COLS>
__device__ __forceinline__ void flush_warp_batch(const IterParams &p, double (*s_acc)[COLS], uint32_t q, uint32_t widx, bool have) {
    if (!have) return;
    double4 nb = make_double4(0, 0, 0, 0);
    bool ok = false;
    if (widx != kNil) {
        const double4 s = ld256(p.src + q);
        nb = ldg256(p.blk_pts + widx);
        ok = accept_pair(nb, s.x, s.y, s.z, p.max_dist);
        if (ok) accumulate_pair(s_acc, threadIdx.x, p.kern, s.x, s.y, s.z, nb.x, nb.y, nb.z);
    }
    if (p.tgt_out) {
        st256(p.tgt_out + q, nb);
        p.matched_out[q] = ok ? 1 : 0;
    }
}

// Adds the `count` partials of each of the 17 sums in a fixed order (partials[k * count + i]): warp w owns sums w, w+W, w+2W, ...;
// lane l adds partials l, l+32, ... of each in order, then a fixed butterfly — the same tree for a given count, so results are
// reproducible.  The loads of all the sums a warp owns are issued together (one round trip per 32 partials instead of one per sum).
// Result in s_sums (shared) and, if given, in `out` (global).  Called by every thread of the block; no barrier inside.
__device__ __forceinline__ void reduce_partials(const double *partials, uint32_t count, double *s_sums, double *out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, kWarps = blockDim.x >> 5;
    constexpr int kMaxOwn = (kSums + 1) / 2;  // >= ceil(kSums / kWarps) for kWarps >= 2 (blocks have at least 64 threads)
    double v[kMaxOwn];
#pragma unroll
    for (int i = 0; i < kMaxOwn; ++i) v[i] = 0.0;
#pragma unroll 2
    for (uint32_t b = lane; b < count; b += 32) {
        double a[kMaxOwn];
#pragma unroll
        for (int i = 0; i < kMaxOwn; ++i) {
            const int k = warp + kWarps * i;
            a[i] = k < kSums ? __ldcg(partials + (size_t)k * count + b) : 0.0;
        }
#pragma unroll
        for (int i = 0; i < kMaxOwn; ++i) v[i] += a[i];
    }
#pragma unroll
    for (int i = 0; i < kMaxOwn; ++i) {
        const int k = warp + kWarps * i;
        if (k < kSums) {  // warp-uniform
            double t = v[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (lane == 0) {
                s_sums[k] = t;
                if (out) out[k] = t;
            }
        }
    }
}

// State a persistent kernel keeps on chip when EVERY block takes the Gauss-Newton step (single rank): the current estimate, the
// iteration count and the stop flag never travel through global memory between iterations.
struct LoopState {
    Pose est;       // transform the next iteration applies to the queries
    Pose T_icp;     // block 0 only: the accumulated estimate
    double norm, est_th;
    double sums[kSums];
    int it, done;
};

// ---------------------------------------------------------------------------------------------
// The part of an iteration that ONE block runs once every partial sum is published: add the `count` partials of each sum in a
// fixed order (p.partials[k * count + i]), exchange the sums with the other ranks (fused peer-memory all-reduce, `tag` = this
// iteration's sequence number), take the Gauss-Newton step.  Called by every thread of that block.
__device__ __forceinline__ void reduce_and_step(const IterParams &p, uint32_t count, Pose &s_est, double &s_norm, unsigned long long tag) {
    __shared__ double s_sums[kSums];  // the reduced sums stay on chip for the solve (st->sums is the copy the host / the exchange reads)
    IcpState *st = p.st;
    __threadfence();
    if (p.dbg && threadIdx.x == 0) p.dbg[kDbg * gridDim.x] = gtime();
    reduce_partials(p.partials, count, s_sums, st->sums);
    __syncthreads();
    if (p.xchg_world > 1) {
        // All-reduce of the 17 sums fused into this kernel: every rank's last block stores its sums, then the launch's tag,
        // into its slot of EVERY peer's exchange buffer (plain stores over NVLink/NVSwitch peer mappings), waits until the
        // slots of all ranks in its own buffer carry the tag, and adds them in rank order — the same values in the same
        // order on every rank, so all replicas take the bit-identical Gauss-Newton step.  Slots alternate by tag parity: a
        // rank can be at most one exchange ahead of a peer.  A peer that never shows up (xchg_timeout_ns, default 30 s: ranks are
        // independent processes whose launches can drift apart by a module load or an allocation) ends the registration.
        const int t = threadIdx.x;
        const size_t slot = ((size_t)(tag & 1ull) * kMaxPeers) * kXchgSlot;
        if (t < p.xchg_world) {
            // payload, then the tag with a RELEASE store at system scope: it orders this thread's own payload stores before the
            // tag as seen by the peer, which is all that is needed — no device-wide fence.sc.sys (MEMBAR.SC.SYS drains every
            // outstanding store of the SM and cost several microseconds per iteration in round 1)
            double *dst = p.xchg_peer[t] + slot + (size_t)p.xchg_rank * kXchgSlot;
#pragma unroll
            for (int k = 0; k < kSums; ++k) asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(dst + k), "d"(s_sums[k]) : "memory");
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(dst + kSums), "l"(tag) : "memory");
        }
        if (t < p.xchg_world) {
            // wait for rank t's tag in this rank's own buffer (ACQUIRE at system scope), then its payload is visible
            const unsigned long long *flag =
                reinterpret_cast<const unsigned long long *>(p.xchg_peer[p.xchg_rank] + slot + (size_t)t * kXchgSlot + kSums);
            const unsigned long long t0 = gtime();
            unsigned long long seen = 0;
            while (true) {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(seen) : "l"(flag) : "memory");
                if (seen == tag) break;
                if (gtime() - t0 > p.xchg_timeout_ns) {
                    st->comm_error = 1;
                    break;
                }
            }
        }
        __syncthreads();
        if (t < kSums) {
            const double *mine = p.xchg_peer[p.xchg_rank] + slot;
            double v = 0;
            for (int r = 0; r < p.xchg_world; ++r) {
                double a;
                asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(a) : "l"(mine + (size_t)r * kXchgSlot + t) : "memory");
                v += a;
            }
            st->sums[t] = v, s_sums[t] = v;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        st->ticket = 0;
        if (p.dbg) p.dbg[kDbg * gridDim.x + 1] = gtime();
        if (p.xchg_world > 1 && st->comm_error) st->done = 1;
    }
    __syncthreads();
    if (p.solve && !(p.xchg_world > 1 && st->comm_error)) icp_step_block(st, s_sums, &s_est, &s_norm, p.dbg ? p.dbg + kDbg * gridDim.x : nullptr);
    if (p.dbg && threadIdx.x == 0) p.dbg[kDbg * gridDim.x + 2] = gtime(), p.dbg[kDbg * gridDim.x + 3] = gridDim.x;
}


// End of one Gauss-Newton iteration for the kernels whose blocks keep running sums: block-reduce the per-thread columns of s_acc
// (COLS columns, a multiple of 32), publish the block's partials, elect the last block to finish, which runs reduce_and_step.
template <int COLS>
__device__ __forceinline__ void finish_iteration(const IterParams &p, double (*s_acc)[COLS], Pose &s_est, double &s_norm, int &s_last,
                                                 unsigned long long tag, double *publish_only = nullptr) {
    static_assert(COLS % 32 == 0, "whole warps of columns");
    IcpState *st = p.st;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, kWarps = blockDim.x >> 5;
    double *partials = publish_only ? publish_only : p.partials;
    // per-block partial sums: threads (fixed tree per sum) -> blocks (fixed order, by the last block to finish)
    __syncthreads();
    if (p.dbg && threadIdx.x == 0) p.dbg[kDbg * blockIdx.x + 3] = gtime();
    for (int k = warp; k < kSums; k += kWarps) {
        double v = 0;
#pragma unroll
        for (int t = 0; t < COLS / 32; ++t) v += s_acc[k][lane + 32 * t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) {
            partials[(size_t)k * gridDim.x + blockIdx.x] = v;  // [sum][block]
            __threadfence();
        }
    }
    if (publish_only) return;  // the caller's grid barrier orders the partials; every block reduces them itself
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return;
    reduce_and_step(p, gridDim.x, s_est, s_norm, tag);
}

// ---------------------------------------------------------------------------------------------
// The search kernel: one launch per Gauss-Newton iteration.
// Light phase — one thread per query: home voxel, then neighbours nearest-bounding-box first with the prune bound
// re-tightened after every voxel, at most `light_probes` neighbour probes.  A query that still has open neighbours after
// that (far from every map point: it has to look at most of its 27 voxels) is put on the block's deferred list.
// Heavy phase — the block's 8 warps share the deferred queries, one warp per query (search_query_warp).
// Queries are dealt to warps in chunks of 32 consecutive points, chunk c to block c % grid, so every block sees a
// cross-section of the scan and the expensive regions (sparse, far from the sensor) spread over all SMs.
#define SAGE_STAMP(i) do { if (p.dbg && pass == 0 && warp == 0) { __syncwarp(); if (lane == 0) p.dbg[kDbg * blockIdx.x + (i)] = gtime(); } } while (0)
template <bool COUNT>
__device__ __forceinline__ void nn_search_iteration(const IterParams &p, const Pose *est_in = nullptr, double *publish_only = nullptr) {
    constexpr int kWarps = kNnThreads / 32;
    // per-thread running sums live in shared memory (s_acc[k][thread]) so that the search loop keeps its registers
    __shared__ double s_acc[kSums][kNnThreads];
    __shared__ Pose s_est;
    __shared__ uint32_t s_cnt[kWarps];
    __shared__ uint32_t s_list[kNnThreads];
    __shared__ Carried s_carry[kNnThreads];
    __shared__ double s_norm;
    __shared__ int s_last;
    IcpState *st = p.st;
    if (est_in == nullptr && p.respect_done && __ldcg(&st->done)) return;
    if (threadIdx.x == 0) s_est = est_in ? *est_in : load_pose_cg(&st->est);
    if (p.dbg && threadIdx.x == 0) p.dbg[kDbg * blockIdx.x] = gtime();
#pragma unroll
    for (int k = 0; k < kSums; ++k) s_acc[k][threadIdx.x] = 0.0;
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double vs = p.voxel_size;
    const float vs32 = p.vs32, th32 = p.th32;
    const float INF = __int_as_float(0x7f800000);
    unsigned long long n_scanned = 0, n_probes = 0, n_exact = 0, n_heavy = 0;

    if (p.all_warp) {
        // small scans (pipeline level: a few thousand queries): give every query a whole warp straight away
        uint32_t bq = 0, bw = kNil;  // lane k keeps the k-th query this warp searched since the last flush
        bool bhave = false;
        int fill = 0;
        for (uint32_t q = blockIdx.x * kWarps + warp; q < p.n; q += gridDim.x * kWarps) {
            double4 s = ld256(p.src + q);
            if (p.apply_est) {  // source <- est * source, in place (core/Registration.cpp:133); every lane computes the same
                const Pose est = s_est;
                double x, y, z;
                pose_act(est, s.x, s.y, s.z, x, y, z);
                s.x = x, s.y = y, s.z = z;
                if (lane == 0) st256(p.src + q, s);
            }
            const uint32_t w = search_query_warp<COUNT>(p, lane, s, Carried{INF, INF, kNil, 0u}, n_scanned, n_probes, n_exact);
            if (lane == fill) bq = q, bw = w, bhave = true;
            if (COUNT && lane == 0) n_heavy += 1;
            if (++fill == 32) {
                __syncwarp();  // lane 0's stores of the transformed points are visible to the lanes that flush them
                flush_warp_batch(p, s_acc, bq, bw, bhave);
                bhave = false, fill = 0;
            }
        }
        __syncwarp();
        flush_warp_batch(p, s_acc, bq, bw, bhave);
    }
    const uint32_t n_chunks = p.all_warp ? 0u : (p.n + 31) / 32, chunks_per_pass = gridDim.x * kWarps;
    const uint32_t passes = (n_chunks + chunks_per_pass - 1) / chunks_per_pass;  // same for every warp of every block
    for (uint32_t pass = 0; pass < passes; ++pass) {
        const uint32_t chunk = (pass * kWarps + warp) * gridDim.x + blockIdx.x;
        const uint32_t wbase = chunk * 32u, q = wbase + lane;
        const bool valid = chunk < n_chunks && q < p.n;
        int kx, ky, kz;
        float bx, by, bz, qlf;
        {
            double sx = 0, sy = 0, sz = 0, sl = 0;
            if (valid) {
                const double4 s = ld256(p.src + q);
                sx = s.x, sy = s.y, sz = s.z, sl = s.w;
                if (p.apply_est) {  // source <- est * source, in place (core/Registration.cpp:133)
                    const Pose est = s_est;
                    pose_act(est, s.x, s.y, s.z, sx, sy, sz);
                    st256(p.src + q, make_double4(sx, sy, sz, sl));
                }
            }
            kx = trunc_div(sx, vs), ky = trunc_div(sy, vs), kz = trunc_div(sz, vs);
            // query relative to its home voxel's origin, and its label, in f32
            bx = hot_offset(sx, kx, vs), by = hot_offset(sy, ky, vs), bz = hot_offset(sz, kz, vs);
            qlf = hot_label(sl);
        }
        // `odd` marks queries the f32 ranking cannot serve
        bool odd = !p.fast_ok || (qlf != qlf) || !(fabsf(bx) <= 2.0f * vs32 && fabsf(by) <= 2.0f * vs32 && fabsf(bz) <= 2.0f * vs32);
        bool heavy = false;
        uint32_t widx = kNil;
        __syncwarp();  // the transformed points of this pass are visible to the whole warp (the exact path re-reads them)
        SAGE_STAMP(4);

        float min1 = INF, min2 = INF;
        uint32_t idx1 = kNil, visited_bits = 0;
        uint32_t hblk = 0, hcnt = 0;
        const bool hfound = valid && !odd && key_in_range(kx, ky, kz) && tbl_find(p.tbl, p.mask, pack_key(kx, ky, kz), hblk, hcnt) && hcnt > 0;
        SAGE_STAMP(5);
        if (hfound) {
            scan_voxel_thread(p.blk_hot, hblk * (uint32_t)p.stride, hcnt, bx, by, bz, qlf, th32, min1, min2, idx1, odd);
            if (COUNT) n_scanned += hcnt;
        }
        SAGE_STAMP(6);
        if (valid && !odd) {
            if (COUNT) n_probes += 1;
            // ---- neighbours, nearest bounding box first, while one can still beat the best so far ----
            float sxm, sxp, sym, syp, szm, szp;
            axis_bounds(bx, kx, vs32, p.box_margin, p.smin32, sxm, sxp);
            axis_bounds(by, ky, vs32, p.box_margin, p.smin32, sym, syp);
            axis_bounds(bz, kz, vs32, p.box_margin, p.smin32, szm, szp);
            uint32_t visited = 1u << 13;  // bit (ox+1)*9 + (oy+1)*3 + (oz+1): the reference's enumeration index
            float bound = prune_bound(p, min1);
            int budget = p.light_probes;
#pragma unroll 1
            while (true) {
                float best_lb = INF;
                int nn = -1;
#pragma unroll
                for (int ox = 0; ox < 3; ++ox) {
                    const float ax = ox == 0 ? sxm : (ox == 2 ? sxp : 0.0f);
                    if (ax > bound) continue;
#pragma unroll
                    for (int oy = 0; oy < 3; ++oy) {
                        const float ay = ax + (oy == 0 ? sym : (oy == 2 ? syp : 0.0f));
                        if (ay > bound) continue;
#pragma unroll
                        for (int oz = 0; oz < 3; ++oz) {
                            const float az = ay + (oz == 0 ? szm : (oz == 2 ? szp : 0.0f));
                            const int id = ox * 9 + oy * 3 + oz;
                            if (az < best_lb && !((visited >> id) & 1u)) best_lb = az, nn = id;
                        }
                    }
                }
                if (nn < 0 || !(best_lb <= bound)) break;
                if (budget-- <= 0) {
                    heavy = true;
                    visited_bits = visited;
                    break;
                }
                visited |= 1u << nn;
                const int ox = nn / 9 - 1, oy = (nn / 3) % 3 - 1, oz = nn % 3 - 1;
                const int nx = kx + ox, ny = ky + oy, nz = kz + oz;
                uint32_t nblk = 0, ncnt = 0;
                if (COUNT) n_probes += 1;
                if (key_in_range(nx, ny, nz) && tbl_find(p.tbl, p.mask, pack_key(nx, ny, nz), nblk, ncnt) && ncnt > 0) {
                    scan_voxel_thread(p.blk_hot, nblk * (uint32_t)p.stride, ncnt, bx - (float)ox * vs32, by - (float)oy * vs32,
                                      bz - (float)oz * vs32, qlf, th32, min1, min2, idx1, odd);
                    if (COUNT) n_scanned += ncnt;
                    bound = prune_bound(p, min1);
                }
            }
            // ---- decide: unique within the error band => idx1 IS the f64 arg-min ----
            if (!heavy && !odd && min1 < INF) {
                if (min2 > band_limit(p, min1))
                    widx = idx1;
                else
                    odd = true;
            }
        }

        SAGE_STAMP(7);
        // ---- queries whose f32 ranking was ambiguous: the whole warp re-ranks them in f64, one at a time ----
        unsigned em = __ballot_sync(0xffffffffu, valid && odd && !heavy);
        while (em) {
            const int l = __ffs(em) - 1;
            em &= em - 1;
            const double4 c = ld256(p.src + wbase + l);
            const uint32_t w = nn_exact<32>(p, 0xffffffffu, lane, c.x, c.y, c.z, c.w, trunc_div(c.x, vs), trunc_div(c.y, vs),
                                            trunc_div(c.z, vs));
            if (lane == l) widx = w;
            if (COUNT && lane == 0) n_exact += 1;
        }

        SAGE_STAMP(8);
        // ---- acceptance, residual, weight and the sums ----
        {
            double4 nb = make_double4(0, 0, 0, 0);
            bool ok = false;
            if (widx != kNil) {
                const double4 s = ld256(p.src + q);  // this thread's own (transformed) point again: cheaper than 8 live registers
                nb = ldg256(p.blk_pts + widx);
                ok = accept_pair(nb, s.x, s.y, s.z, p.max_dist);
                if (ok) accumulate_pair(s_acc, threadIdx.x, p.kern, s.x, s.y, s.z, nb.x, nb.y, nb.z);
            }
            if (p.tgt_out && valid && !heavy) {
                st256(p.tgt_out + q, nb);
                p.matched_out[q] = ok ? 1 : 0;
            }
        }

        SAGE_STAMP(9);
        // ---- deferred queries of this pass: deterministic compaction (warp order, lane order), then one warp per query ----
        const unsigned hm = __ballot_sync(0xffffffffu, heavy);
        if (lane == 0) s_cnt[warp] = __popc(hm);
        if (p.dbg && threadIdx.x == 0 && pass == 0) p.dbg[kDbg * blockIdx.x + 1] = gtime();  // warp 0 done with its light phase
        __syncthreads();
        uint32_t before = 0, total = 0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            before += w < warp ? s_cnt[w] : 0u;
            total += s_cnt[w];
        }
        if (heavy) {
            const uint32_t slot = before + __popc(hm & ((1u << lane) - 1u));
            s_list[slot] = q;
            // a query that met a record the f32 ranking cannot serve (`odd`) starts over in the warp phase, which will meet it again
            s_carry[slot] = odd ? Carried{INF, INF, kNil, 0u} : Carried{min1, min2, idx1, visited_bits};
        }
        __syncthreads();
        if (p.dbg && threadIdx.x == 0 && pass == 0) p.dbg[kDbg * blockIdx.x + 2] = gtime();  // whole block done with the light phase
        {
            uint32_t bq = 0, bw = kNil;
            bool bhave = false;
            int fill = 0;
            for (uint32_t i = warp; i < total; i += kWarps) {
                const uint32_t hq = s_list[i];
                const double4 s = ld256(p.src + hq);  // transformed above (same block, ordered by the barrier)
                const uint32_t w = search_query_warp<COUNT>(p, lane, s, s_carry[i], n_scanned, n_probes, n_exact);
                if (lane == fill) bq = hq, bw = w, bhave = true;
                if (COUNT && lane == 0) n_heavy += 1;
                if (++fill == 32) {
                    flush_warp_batch(p, s_acc, bq, bw, bhave);
                    bhave = false, fill = 0;
                }
            }
            flush_warp_batch(p, s_acc, bq, bw, bhave);
        }
        if (pass + 1 < passes) __syncthreads();  // s_cnt / s_list are reused by the next pass
    }

    if (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_scanned += __shfl_xor_sync(0xffffffffu, n_scanned, o);
            n_probes += __shfl_xor_sync(0xffffffffu, n_probes, o);
            n_exact += __shfl_xor_sync(0xffffffffu, n_exact, o);
            n_heavy += __shfl_xor_sync(0xffffffffu, n_heavy, o);
        }
        if (lane == 0) {
            atomicAdd(&st->stat_scanned, n_scanned);
            atomicAdd(&st->stat_probes, n_probes);
            atomicAdd(&st->stat_exact, n_exact);
            atomicAdd(&st->stat_heavy, n_heavy);
        }
    }
    finish_iteration<kNnThreads>(p, s_acc, s_est, s_norm, s_last, p.xchg_tag, publish_only);
}

template <bool COUNT>
__global__ void __launch_bounds__(kNnThreads, SAGE_LIGHT_MINB) nn_search_kernel(IterParams p) {
    nn_search_iteration<COUNT>(p);
}
// The whole Gauss-Newton loop of one registration in ONE cooperative launch: every block runs the iteration above, the grid
// meets at a barrier (the last block has solved and updated IcpState by then), and the loop ends when `done` is set — no
// launch gap, no host poll between iterations.  Used for the small scans of the pipeline level, where an iteration is ~20 us
// of work and the gaps were as long as the work (profiles/r01g_streaming.md).  Launched with cudaLaunchCooperativeKernel, which
// refuses a grid that is not co-resident, so the barrier cannot deadlock.
// End of an iteration when every block takes the step (persistent kernels, single rank): after the grid barrier that made the
// partials of all blocks visible, each block adds them in the same fixed order, solves and exponentiates — identical bits in every
// block — and keeps the new estimate on chip.  Block 0 also keeps the books the host reads at the end.  Compared with electing a last
// block this takes the ticket, the serial reduction, the pose product and the reload of the estimate off every iteration's path
// and saves the second grid-wide wait.
__device__ __forceinline__ void loop_step_everywhere(const IterParams &p, LoopState &ls, const double *partials, uint32_t count, int max_iterations) {
    reduce_partials(partials, count, ls.sums, nullptr);
    __syncthreads();
    icp_step_local(ls.sums, &ls.est, &ls.norm, ls.est_th);
    if (blockIdx.x == 0 && threadIdx.x == 0) ls.T_icp = pose_mul(ls.est, ls.T_icp);
    __syncthreads();
    if (threadIdx.x == 0) {
        ls.it += 1;
        ls.done = (ls.norm < ls.est_th || ls.it >= max_iterations) ? 1 : 0;
    }
    __syncthreads();
}
__device__ __forceinline__ void loop_state_init(const IterParams &p, LoopState &ls) {
    if (threadIdx.x == 0) {
        ls.est = load_pose_cg(&p.st->est);
        ls.T_icp = pose_identity();
        ls.it = 0, ls.done = __ldcg(&p.st->done), ls.norm = 0, ls.est_th = p.st->est_th;
    }
    __syncthreads();
}
__device__ __forceinline__ void loop_state_commit(const IterParams &p, const LoopState &ls) {  // block 0, after the loop
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    IcpState *st = p.st;
    st->est = ls.est, st->T_icp = ls.T_icp, st->iter = ls.it, st->last_norm = ls.norm;
    for (int k = 0; k < kSums; ++k) st->sums[k] = ls.sums[k];
    if (ls.it > 0) st->result = pose_mul(ls.T_icp, st->guess);  // T_icp * guess, core/Registration.cpp:140
    st->done = 1;
}

// MINB: resident blocks per SM the register allocation aims at.  A pipeline-level cloud (a few hundred to ~2 000 queries, a warp
// each) fills at most two blocks per SM, so it runs the instantiation with 128 registers per thread: the 64-register one spills
// ~4 KB per thread (ptxas -v), and local-memory round trips sit on the dependent chain of every query.
template <int MINB>
__global__ void __launch_bounds__(kNnThreads, MINB) nn_search_persistent_kernel(IterParams p, int max_iterations, int first_apply) {
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    if (p.step_everywhere) {
        __shared__ LoopState ls;
        loop_state_init(p, ls);
        while (!ls.done) {
            p.apply_est = ls.it > 0 ? 1 : first_apply;  // first_apply = 0: the caller has already applied the initial guess
            double *partials = p.partials + (size_t)(ls.it & 1) * kSums * gridDim.x;  // two buffers: a block may be one iteration ahead
            nn_search_iteration<false>(p, &ls.est, partials);
            grid.sync();
            loop_step_everywhere(p, ls, partials, gridDim.x, max_iterations);
        }
        loop_state_commit(p, ls);
        return;
    }
    for (int i = 0; i < max_iterations; ++i) {
        p.apply_est = i > 0 ? 1 : first_apply;  // first_apply = 0: the caller has already applied the initial guess (tile_prepare)
        nn_search_iteration<false>(p);
        grid.sync();  // block barrier + grid barrier + fence: the new estimate and `done` are visible to every block
        if (*reinterpret_cast<volatile int *>(&p.st->done)) break;
    }
}

}  // namespace sage
#include "search_tile.cuh"
namespace sage {

// Neighbourhood statistics for the algorithmic-bytes figure: per query, how many of the 27 voxels exist
// and how many points they hold.  One thread per (query, voxel).
__global__ void nn_stats_kernel(const TblEntry *tbl, uint32_t mask, const double4 *src, uint32_t n, double vs, IcpState *st) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t q = t / 32, pr = t % 32;
    unsigned long long occ = 0, cand = 0;
    if (q < n && pr < 27) {
        const double4 s = src[q];
        const int nx = trunc_div(s.x, vs) + (int)pr / 9 - 1, ny = trunc_div(s.y, vs) + ((int)pr / 3) % 3 - 1, nz = trunc_div(s.z, vs) + (int)pr % 3 - 1;
        uint32_t b, c;
        if (key_in_range(nx, ny, nz) && tbl_find(tbl, mask, pack_key(nx, ny, nz), b, c)) occ = 1, cand = c;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        occ += __shfl_xor_sync(0xffffffffu, occ, o);
        cand += __shfl_xor_sync(0xffffffffu, cand, o);
    }
    if ((threadIdx.x & 31) == 0 && occ) {
        atomicAdd(&st->stat_occupied, occ);
        atomicAdd(&st->stat_candidates, cand);
    }
}

// ---------------------------------------------------------------------------------------------
// host side

void VoxelMapGPU::profile_enable(bool on) {
    profile_ = on;
    prof_used_ = 0;
}

// launches = Gauss-Newton iterations timed (a persistent launch counts every iteration it ran), ms = their total device time
void VoxelMapGPU::profile_read(long long *launches, double *ms, long long *kernel_launches) {
    set_device();
    SAGE_CUDA(cudaStreamSynchronize(stream_));
    double total = 0;
    long long iters = 0;
    for (size_t i = 0; i < prof_used_; ++i) {
        float t = 0;
        SAGE_CUDA(cudaEventElapsedTime(&t, prof_events_[i].first, prof_events_[i].second));
        total += t;
        iters += prof_iters_[i];
    }
    if (launches) *launches = iters;
    if (kernel_launches) *kernel_launches = (long long)prof_used_;
    if (ms) *ms = total;
    prof_used_ = 0;
}

void VoxelMapGPU::prof_begin() {
    if (!profile_) return;
    if (prof_used_ == prof_events_.size()) {
        cudaEvent_t a, b;
        SAGE_CUDA(cudaEventCreate(&a));
        SAGE_CUDA(cudaEventCreate(&b));
        prof_events_.emplace_back(a, b);
        prof_iters_.push_back(1);
    }
    SAGE_CUDA(cudaEventRecord(prof_events_[prof_used_].first, stream_));
}
void VoxelMapGPU::prof_end(int iterations) {
    if (!profile_) return;
    SAGE_CUDA(cudaEventRecord(prof_events_[prof_used_].second, stream_));
    prof_iters_[prof_used_] = iterations;
    ++prof_used_;
}

static const void *tile_kernel_ptr(int minb, bool persistent) {
    if (persistent) {
        switch (minb) {
            case 4: return (const void *)nn_tile_persistent_kernel<4>;
            case 5: return (const void *)nn_tile_persistent_kernel<5>;
            case 6: return (const void *)nn_tile_persistent_kernel<6>;
            default: return (const void *)nn_tile_persistent_kernel<8>;
        }
    }
    switch (minb) {
        case 4: return (const void *)nn_tile_kernel<4>;
        case 5: return (const void *)nn_tile_kernel<5>;
        case 6: return (const void *)nn_tile_kernel<6>;
        default: return (const void *)nn_tile_kernel<8>;
    }
}

void preload_registration_kernels() {
    const void *ks[] = {(const void *)nn_search_kernel<false>, (const void *)nn_search_kernel<true>,
                        (const void *)nn_search_persistent_kernel<1>, (const void *)nn_search_persistent_kernel<2>,
                        (const void *)nn_search_persistent_kernel<SAGE_LIGHT_MINB>,
                        (const void *)nn_tile_kernel<4>, (const void *)nn_tile_kernel<5>, (const void *)nn_tile_kernel<6>, (const void *)nn_tile_kernel<8>,
                        (const void *)nn_tile_kernel<4, true>,
                        (const void *)nn_tile_persistent_kernel<4>, (const void *)nn_tile_persistent_kernel<5>,
                        (const void *)nn_tile_persistent_kernel<6>, (const void *)nn_tile_persistent_kernel<8>,
                        (const void *)nn_stats_kernel, (const void *)icp_init_kernel, (const void *)icp_solve_kernel};
    for (const void *k : ks) preload_kernel(k);
}

static long env_long(const char *name, long dflt) {
    const char *e = getenv(name);
    return e ? atol(e) : dflt;
}

// one-time launch configuration of the search kernels (grids = co-resident blocks, tuning knobs from the environment)
void VoxelMapGPU::init_search_config() {
    if (nn_grid_ != 0) return;
    int per_sm = 0;
    SAGE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nn_search_kernel<false>, kNnThreads, 0));
    nn_grid_ = sm_count_ * (per_sm > 0 ? per_sm : 1);
    if (const char *e = getenv("SAGE_LIGHT_PROBES")) light_probes_ = atoi(e);  // tuning knobs
    all_warp_max_ = (size_t)nn_grid_ * (kNnThreads / 32) * 3 / 4;
    all_warp_max_ = (size_t)env_long("SAGE_ALL_WARP_MAX", (long)all_warp_max_);
    if (getenv("SAGE_NO_ALL_WARP")) all_warp_max_ = 0;
    int coop = 0;
    SAGE_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device_));
    int per_sm_p = 0;
    SAGE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_p, nn_search_persistent_kernel<SAGE_LIGHT_MINB>, kNnThreads, 0));
    persistent_grid_ = sm_count_ * per_sm_p;
    if (persistent_grid_ > nn_grid_) persistent_grid_ = nn_grid_;
    {
        int per_sm_w = 0;
        SAGE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_w, nn_search_persistent_kernel<2>, kNnThreads, 0));
        const long wide = env_long("SAGE_SMALL_WIDE", 2);  // 0: always the 64-register instantiation, 1: 128 where it fits, 2: 255 / 128
        persistent_grid_wide_ = wide != 0 ? sm_count_ * per_sm_w : 0;
        SAGE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_w, nn_search_persistent_kernel<1>, kNnThreads, 0));
        persistent_grid_widest_ = wide >= 2 ? sm_count_ * (per_sm_w > 1 ? 1 : per_sm_w) : 0;
    }
    if (persistent_grid_ < 1) coop = 0;
    coop_ok_ = coop != 0;
    persistent_max_ = coop ? 20000 : 0;  // scans up to this many queries run their GN loop in one cooperative launch
    persistent_max_ = coop ? (size_t)env_long("SAGE_PERSISTENT_MAX", (long)persistent_max_) : 0;
    // tile search: staging area (dynamic shared memory) + co-resident grid
    static_assert(kTileThreads == 128, "tile_sort.cu cuts units of 128 queries");
    tile_stage_cap_ = (uint32_t)env_long("SAGE_TILE_STAGE", 1536);  // records of 16 bytes: 24 KB
    tile_minb_ = (int)env_long("SAGE_TILE_MINB", 4);  // which instantiation: resident blocks per SM the register allocation aims at
    if (tile_minb_ < 4) tile_minb_ = 4;
    if (tile_minb_ > 6) tile_minb_ = 8;  // 4, 5, 6 or 8 blocks per SM: 128, 96, 80 or 64 registers per thread
    const size_t smem = (size_t)tile_stage_cap_ * sizeof(float4);
    const void *k1 = tile_kernel_ptr(tile_minb_, false), *k2 = tile_kernel_ptr(tile_minb_, true);
    SAGE_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SAGE_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SAGE_CUDA(cudaFuncSetAttribute(nn_tile_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm_t = 0, per_sm_tp = 0;
    SAGE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_t, k1, kTileThreads, smem));
    SAGE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_tp, k2, kTileThreads, smem));
    int per_sm_tile = per_sm_t < per_sm_tp ? per_sm_t : per_sm_tp;
    const int want = (int)env_long("SAGE_TILE_BLOCKS", per_sm_tile);
    if (want >= 1 && want < per_sm_tile) per_sm_tile = want;
    tile_grid_ = sm_count_ * per_sm_tile;
    tile_min_ = tile_grid_ > 0 ? (size_t)env_long("SAGE_TILE_MIN", 12288) : 0;
    if (tile_grid_ > 0 && tile_min_ < 1) tile_min_ = 1;
    if (env_long("SAGE_TILE", 1) == 0) tile_min_ = 0;
    tile_persistent_ = coop_ok_ && env_long("SAGE_TILE_PERSISTENT", 1) != 0;
    tile_by_size_ = env_long("SAGE_TILE_BY_SIZE", 1) != 0;
    step_everywhere_ = (int)env_long("SAGE_STEP_EVERYWHERE", 1);  // 0 never, 1 where it pays (tile search; small scans on the wide instantiations), 2 wherever possible
    tile_fill_ = (size_t)env_long("SAGE_TILE_FILL", 1);  // 0: the tile search never declines a thinly spread query set
    tile_graph_ = env_long("SAGE_TILE_GRAPH", 1) != 0;   // sort + unit list replayed as a captured CUDA graph (tile_sort.cu)
    {
        const long t = env_long("SAGE_XCHG_TIMEOUT_S", 30);
        xchg_timeout_ns_ = (unsigned long long)(t < 1 ? 1 : t) * 1000000000ull;
    }
    partials_.ensure((size_t)2 * kSums * (nn_grid_ > tile_grid_ ? nn_grid_ : tile_grid_));  // two buffers (iteration parity)
    // The first COOPERATIVE launch of a kernel is slow even when its code is loaded (10-35 ms on B200: the driver sets the launch
    // up per function), and which instantiation a registration needs depends on its query count — on a live drive that was a stall
    // at the first frame with more than 1 184 queries.  So every persistent kernel is launched once here, on one block, with zero
    // iterations (they leave at once and touch nothing).
    if (coop_ok_) {
        IterParams p = {};
        int zero = 0;
        void *args3[] = {&p, &zero, &zero};
        const void *small[] = {(const void *)nn_search_persistent_kernel<1>, (const void *)nn_search_persistent_kernel<2>,
                               (const void *)nn_search_persistent_kernel<SAGE_LIGHT_MINB>};
        for (const void *k : small) SAGE_CUDA(cudaLaunchCooperativeKernel(k, dim3(1), dim3(kNnThreads), args3, 0, stream_));
        void *args2[] = {&p, &zero};
        SAGE_CUDA(cudaLaunchCooperativeKernel(k2, dim3(1), dim3(kTileThreads), args2, smem, stream_));
        SAGE_CUDA(cudaStreamSynchronize(stream_));
    }
}

// the rule by which nn_tile_iteration declines a query set (search_tile.cuh), for the paths that decide on the host
bool VoxelMapGPU::tile_units_too_thin(uint32_t n_units, size_t n) const {
    const unsigned long long rounds = ((unsigned long long)n_units + tile_grid_ - 1) / (unsigned long long)tile_grid_;
    return rounds * 15000ull + 12000ull > 30000ull + (unsigned long long)n * 48ull / 100ull;
}

void VoxelMapGPU::fill_params(IterParams &p, double4 *src, size_t n, double max_dist, double kernel, double sem_th, int mode, double4 *tgt_out,
                              uint8_t *matched_out) {
    p.tbl = tbl_.p, p.mask = tbl_cap_ - 1, p.blk_pts = blk_pts_.p, p.blk_hot = blk_hot_.p, p.stride = stride_, p.voxel_size = voxel_size_;
    p.src = src, p.n = (uint32_t)n;
    p.max_dist = max_dist, p.kern = kernel, p.sem_th = sem_th;
    // f32 error model (derivation in DESIGN.md §4): per-axis error of (record - query) <= 8 u vs, u = 2^-24, so
    // |D32 - D| <= 28 u vs sqrt(D) + 3.1 u D and the metric (D or D * th) adds 2.1 u relative; constants doubled.
    {
        const double u = 5.9604644775390625e-8, vsd = voxel_size_;
        const bool ok = sem_th > 0.0 && sem_th < 1e30 && vsd > 1e-6 && vsd < 1e6;
        const double smin = ok ? (sem_th < 1.0 ? sem_th : 1.0) : 1.0, smax = ok ? (sem_th > 1.0 ? sem_th : 1.0) : 1.0;
        p.fast_ok = ok ? 1 : 0;
        p.vs32 = (float)vsd, p.th32 = (float)sem_th;
        p.smin32 = (float)(smin * (1.0 - 1e-5)), p.inv_smin32 = (float)((1.0 + 1e-5) / smin);
        p.err_scale = (float)(smax * (1.0 + 1e-5));
        p.err_a = (float)(64.0 * u * vsd), p.err_b = (float)(12.0 * u), p.err_c = (float)(1e-11 * vsd * vsd);
        p.box_margin = (float)(32.0 * u * vsd);
    }
    p.st = icp_.p, p.partials = partials_.p, p.tgt_out = tgt_out, p.matched_out = matched_out;
    p.dbg = nullptr;
    p.light_probes = 8, p.all_warp = 0, p.step_everywhere = 0;
    p.apply_est = (mode == 0), p.respect_done = (mode == 0);
    p.solve = (mode == 0 && comm_ == nullptr);  // NCCL path: all-reduce and solve are separate launches
    p.xchg_world = 1, p.xchg_rank = 0, p.xchg_tag = 0, p.xchg_timeout_ns = xchg_timeout_ns_;
    for (int k = 0; k < kMaxPeers; ++k) p.xchg_peer[k] = nullptr;
    if (mode == 0 && peer_world_ > 1) {  // fused peer-memory all-reduce: the search kernel does everything
        p.solve = 1;
        p.xchg_world = peer_world_, p.xchg_rank = peer_rank_, p.xchg_tag = xchg_tag_ + 1;  // + iteration index (callers)
        for (int k = 0; k < peer_world_; ++k) p.xchg_peer[k] = peer_buf_[k];
    }
    p.tile_units = tile_units_.p, p.tile_n_units = tile_nunits_.p, p.tile_perm = tile_vals_[1].p, p.tile_order = tile_by_size_ ? tile_order_.p : nullptr, p.tile_stage_cap = tile_stage_cap_;
    p.tile_unit_part = tile_unit_part_.p, p.tile_group_cnt = tile_group_cnt_.p;
    p.tile_ctl = tile_ctl_.p, p.tile_fill = 0;
}

// mode 0: ICP iteration (apply est, solve on device when single rank); mode 1: correspondences/sums of the points
// as given; mode 2: as mode 1 with the search-work counters on.
void VoxelMapGPU::launch_iteration(double4 *src, size_t n, double max_dist, double kernel, double sem_th, int mode,
                                   double4 *tgt_out, uint8_t *matched_out, int persistent_iters, int iter_index, bool pre_transformed) {
    init_search_config();
    if (mode != 0 && tile_min_ > 0 && n >= tile_min_) {
        // correspondences / sums / work counters of the points as given, through the tile search: gather them in cell order
        // (no transform), one nn_tile_kernel launch, results scattered back through the permutation
        src_.ensure(n);
        tile_prepare(src, n, pose_identity(), false);
        tile_nunits_pin_.ensure(1);
        SAGE_CUDA(cudaMemcpyAsync(tile_nunits_pin_.p, tile_nunits_.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream_));
        SAGE_CUDA(cudaStreamSynchronize(stream_));
        last_units_ = *tile_nunits_pin_.p;
    }
    if (mode != 0 && tile_min_ > 0 && n >= tile_min_ && !(tile_fill_ && tile_units_too_thin(last_units_, n))) {  // same rule as the kernel's
        IterParams p;
        fill_params(p, src_.p, n, max_dist, kernel, sem_th, mode, tgt_out, matched_out);
        const size_t smem = (size_t)tile_stage_cap_ * sizeof(float4);
        if (mode == 2)
            SAGE_LAUNCH((nn_tile_kernel<4, true>), tile_grid_, kTileThreads, smem, stream_, p, 1);
        else {
            int first = 1;
            void *args[] = {&p, &first};
            SAGE_CUDA(cudaLaunchKernel(tile_kernel_ptr(tile_minb_, false), dim3(tile_grid_), dim3(kTileThreads), args, smem, stream_));
            g_launches.fetch_add(1, std::memory_order_relaxed);
        }
        return;
    }
    // chunks of 32 consecutive queries are dealt round-robin to the blocks; a full grid (a multiple of the SM count) once
    // there is a chunk for every block, fewer blocks for small scans
    // small scans: one warp per query (all_warp) as long as that is at most two queries per resident warp
    const bool all_warp = n <= all_warp_max_;
    uint32_t grid = all_warp ? (uint32_t)((n + kNnThreads / 32 - 1) / (kNnThreads / 32)) : (uint32_t)((n + 31) / 32);
    grid = grid < 1 ? 1 : (grid > (uint32_t)nn_grid_ ? (uint32_t)nn_grid_ : grid);
    if (persistent_iters > 0 && grid > (uint32_t)persistent_grid_) grid = (uint32_t)persistent_grid_;  // must be co-resident

    IterParams p;
    fill_params(p, src, n, max_dist, kernel, sem_th, mode, tgt_out, matched_out);
    p.dbg = (mode == 0 && dbg_on_) ? dbg_.p : nullptr;
    // Neighbour probes a query may spend in the thread-per-query phase before it is deferred to the warp phase: the fewer
    // queries there are, the more idle warps the deferred phase finds, so the earlier it pays to hand a query over
    // (measured on B200, profiles/r01g_nn_search_kernel_ncu.md: best of 1/2/3/5/8 at each size).
    const int auto_probes = n <= 20000 ? 1 : (n <= 90000 ? 3 : 8);
    p.light_probes = light_probes_ >= 0 ? light_probes_ : auto_probes, p.all_warp = all_warp ? 1 : 0;
    p.xchg_tag += (unsigned long long)iter_index;  // exchange number = exchanges completed so far + 1 + iteration
    if (pre_transformed && iter_index == 0) p.apply_est = 0;  // src already carries the initial guess (sorted by tile_prepare)

    const bool prof = mode == 0;
    if (prof) prof_begin();
    if (persistent_iters > 0) {
        int first_apply = pre_transformed ? 0 : 1;
        // every block takes the step itself (no elected last block, loop state on chip): with the wide instantiations it pays —
        // 16.2 vs 19.2 us per iteration at 700 queries, 23.9 vs 24.8 at 2 100, 29.1 vs 30.0 at 5 000 — but not on the full grid of
        // the 64-register one (12 000 queries, 375 blocks each re-reading 375 x 17 partials: 58.0 vs 53.2; profiles/
        // r02w_small_scans.md).  SAGE_STEP_EVERYWHERE=0 keeps the elected block, =2 forces every block.
        p.step_everywhere = (peer_world_ <= 1 && comm_ == nullptr &&
                             (step_everywhere_ == 2 || (step_everywhere_ == 1 && grid <= (uint32_t)persistent_grid_wide_))) ? 1 : 0;
        void *args[] = {&p, &persistent_iters, &first_apply};
        const void *kernel = grid <= (uint32_t)persistent_grid_widest_ ? (const void *)nn_search_persistent_kernel<1>
                             : grid <= (uint32_t)persistent_grid_wide_ ? (const void *)nn_search_persistent_kernel<2>
                                                                       : (const void *)nn_search_persistent_kernel<SAGE_LIGHT_MINB>;
        SAGE_CUDA(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(kNnThreads), args, 0, stream_));
        g_launches.fetch_add(1, std::memory_order_relaxed);
    } else if (mode == 2) {
        SAGE_LAUNCH(nn_search_kernel<true>, grid, kNnThreads, 0, stream_, p);
    } else {
        SAGE_LAUNCH(nn_search_kernel<false>, grid, kNnThreads, 0, stream_, p);
    }
    if (prof) prof_end(1);
    if (mode == 0 && comm_ != nullptr && peer_world_ <= 1) {
        nccl_allreduce_sum_f64(comm_, icp_.p->sums, kSums, stream_);
        SAGE_LAUNCH(icp_solve_kernel, 1, 64, 0, stream_, icp_.p);
    }
}

// One Gauss-Newton iteration (persistent_iters == 0) or the whole loop (cooperative launch) of the tile search over src_, which
// tile_prepare() has sorted, transformed by the initial guess and cut into units.
void VoxelMapGPU::launch_tile(size_t n, double max_dist, double kernel, double sem_th, int iter_index, int persistent_iters) {
    IterParams p;
    fill_params(p, src_.p, n, max_dist, kernel, sem_th, 0, nullptr, nullptr);
    p.dbg = dbg_on_ ? dbg_.p : nullptr;
    p.tile_fill = comm_ == nullptr ? (uint32_t)tile_fill_ : 0u;  // NCCL variant: decided on the host (register_frame_dev)
    const size_t smem = (size_t)tile_stage_cap_ * sizeof(float4);
    prof_begin();
    if (persistent_iters > 0) {
        // every block takes the step itself after ONE grid barrier, and warp 0 publishes a unit while the other warps start the next:
        // 69.0 vs 72.7 us per iteration at 120 k queries, 60.8 vs 61.2 at 60 k, 50.0 vs 50.6 at 15 k (visit Z, same box; before the
        // publish moved off the block's path it had lost below 65 536 queries)
        p.step_everywhere = (peer_world_ <= 1 && comm_ == nullptr && step_everywhere_ >= 1) ? 1 : 0;
        void *args[] = {&p, &persistent_iters};
        SAGE_CUDA(cudaLaunchCooperativeKernel(tile_kernel_ptr(tile_minb_, true), dim3(tile_grid_), dim3(kTileThreads), args, smem, stream_));
        g_launches.fetch_add(1, std::memory_order_relaxed);
    } else {
        p.xchg_tag += (unsigned long long)iter_index;
        int first = iter_index == 0 ? 1 : 0;
        void *args[] = {&p, &first};
        SAGE_CUDA(cudaLaunchKernel(tile_kernel_ptr(tile_minb_, false), dim3(tile_grid_), dim3(kTileThreads), args, smem, stream_));
        g_launches.fetch_add(1, std::memory_order_relaxed);
    }
    prof_end(1);
    if (comm_ != nullptr && peer_world_ <= 1) {
        nccl_allreduce_sum_f64(comm_, icp_.p->sums, kSums, stream_);
        SAGE_LAUNCH(icp_solve_kernel, 1, 64, 0, stream_, icp_.p);
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
    // SAGE_TRACE_SLOW=<ms>: a registration that takes longer says where the host spent the time (development aid)
    static const double trace_slow_ms = getenv("SAGE_TRACE_SLOW") ? atof(getenv("SAGE_TRACE_SLOW")) : 0.0;
    const auto tr0 = std::chrono::steady_clock::now();
    icp_.ensure(1);
    icp_pin_.ensure(1);
    src_.ensure(n ? n : 1);
    init_search_config();
    const auto tr1 = std::chrono::steady_clock::now();
    // large scans: sort the queries by cell once, then the tile search (search_tile.cuh); the NCCL variant keeps its separate
    // all-reduce + solve launches, so it cannot run the loop in one launch
    bool tile = tile_min_ > 0 && n >= tile_min_;
    bool sorted = false;
    if (tile) {
        tile_prepare(frame, n, guess, true, true, max_iters, est_th);  // ... and the loop state initialised (icp_state_init)
        sorted = true;  // src_ = the queries in cell order, initial guess applied
        // A query set spread thinly over the map is declined by the kernel itself (`declined` below, search_tile.cuh).  Only the NCCL
        // variant decides here, with a read-back of the unit count: its all-reduce launches must match on every rank.
        if (comm_ != nullptr) {
            tile_nunits_pin_.ensure(1);
            SAGE_CUDA(cudaMemcpyAsync(tile_nunits_pin_.p, tile_nunits_.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream_));
            SAGE_CUDA(cudaStreamSynchronize(stream_));
            last_units_ = *tile_nunits_pin_.p;
            if (tile_fill_ && tile_units_too_thin(last_units_, n)) tile = false;
        }
    } else if (n) {
        SAGE_CUDA(cudaMemcpyAsync(src_.p, frame, n * sizeof(double4), cudaMemcpyDeviceToDevice, stream_));
    }
    if (!sorted) SAGE_LAUNCH(icp_init_kernel, 1, 1, 0, stream_, icp_.p, guess, max_iters, est_th);
    // iterations are launched in batches (kernels of a finished registration return at once) and `done` is polled between
    // batches; the first batch is sized from the previous registration so that the common case needs one round trip
    int launched = 0;
    bool persistent = false;
    const size_t prof_mark = prof_used_;
    if (tile && tile_persistent_ && comm_ == nullptr && !dbg_on_) {
        launch_tile(n, max_dist, kernel, sem_th, 0, max_iters);
        persistent = true;
    } else if (!tile && n <= persistent_max_ && comm_ == nullptr && peer_world_ <= 1 && !dbg_on_) {
        // small scans, single rank: the whole loop in one cooperative launch (nn_search_persistent_kernel)
        launch_iteration(src_.p, n, max_dist, kernel, sem_th, 0, nullptr, nullptr, max_iters, 0, sorted);
        persistent = true;
    }
    const auto tr2 = std::chrono::steady_clock::now();
    if (persistent) {
        SAGE_CUDA(cudaMemcpyAsync(icp_pin_.p, icp_.p, sizeof(IcpState), cudaMemcpyDeviceToHost, stream_));
        SAGE_CUDA(cudaStreamSynchronize(stream_));
        launched = max_iters;
        if (profile_ && prof_used_ > 0) prof_iters_[prof_used_ - 1] = icp_pin_.p->iter > 0 ? icp_pin_.p->iter : 1;
    }
    while (launched < max_iters) {
        int batch = launched == 0 ? (last_iters_ + 2 > 8 ? last_iters_ + 2 : 8) : 8;
        if (batch > 48) batch = 48;
        if (batch > max_iters - launched) batch = max_iters - launched;
        for (int b = 0; b < batch; ++b) {
            if (tile)
                launch_tile(n, max_dist, kernel, sem_th, launched + b, 0);
            else
                launch_iteration(src_.p, n, max_dist, kernel, sem_th, 0, nullptr, nullptr, 0, launched + b, sorted);
        }
        launched += batch;
        SAGE_CUDA(cudaMemcpyAsync(icp_pin_.p, icp_.p, sizeof(IcpState), cudaMemcpyDeviceToHost, stream_));
        SAGE_CUDA(cudaStreamSynchronize(stream_));
        if (icp_pin_.p->done) break;
    }
    if (tile && icp_pin_.p->declined) {
        // The tile kernel found its units too thinly filled (BASELINE configs[4]'s uniform queries, a voxel-downsampled cloud) and
        // did nothing: the per-query kernel takes the sorted array as it is.  No exchange has happened yet, so the ranks stay in step
        // whatever each of them decides.
        tile = false;
        prof_used_ = prof_mark;  // the declined launches did no work: they are not iterations
        SAGE_LAUNCH(icp_init_kernel, 1, 1, 0, stream_, icp_.p, guess, max_iters, est_th);
        launched = 0;
        while (launched < max_iters) {
            int batch = launched == 0 ? (last_iters_ + 2 > 8 ? last_iters_ + 2 : 8) : 8;
            if (batch > 48) batch = 48;
            if (batch > max_iters - launched) batch = max_iters - launched;
            for (int b = 0; b < batch; ++b) launch_iteration(src_.p, n, max_dist, kernel, sem_th, 0, nullptr, nullptr, 0, launched + b, true);
            launched += batch;
            SAGE_CUDA(cudaMemcpyAsync(icp_pin_.p, icp_.p, sizeof(IcpState), cudaMemcpyDeviceToHost, stream_));
            SAGE_CUDA(cudaStreamSynchronize(stream_));
            if (icp_pin_.p->done) break;
        }
    }
    // exchanges completed: every rank ran the same iterations (lock-step), whichever launch mode each of them chose
    if (peer_world_ > 1) xchg_tag_ += (unsigned long long)icp_pin_.p->iter;
    if (icp_pin_.p->comm_error) {
        // the exchange sequence numbers of the ranks no longer agree: drop the mappings so that the next registration fails
        // loudly ("not attached") instead of deadlocking; every rank has to attach again (sage_map_comm_peer_attach)
        peer_detach();
        throw CudaError("peer exchange timed out: a rank of the sharded registration did not arrive; the peer communicator was detached, "
                        "re-attach on every rank (SAGE_XCHG_TIMEOUT_S sets the wait, default 30 s)");
    }
    if (trace_slow_ms > 0) {
        const auto tr3 = std::chrono::steady_clock::now();
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
        if (ms(tr0, tr3) > trace_slow_ms)
            fprintf(stderr, "[sage] slow registration: %zu queries, %d iterations, %.2f ms = buffers + configuration %.2f, enqueue %.2f, wait for the device %.2f\n",
                    n, icp_pin_.p->iter, ms(tr0, tr3), ms(tr0, tr1), ms(tr1, tr2), ms(tr2, tr3));
    }
    pose_out = icp_pin_.p->result;
    last_iters_ = icp_pin_.p->iter;
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
    SAGE_LAUNCH(nn_stats_kernel, (unsigned)((n * 32 + 255) / 256), 256, 0, stream_, tbl_.p, tbl_cap_ - 1, d, (uint32_t)n, voxel_size_, icp_.p);
    SAGE_CUDA(cudaMemcpyAsync(icp_pin_.p, icp_.p, sizeof(IcpState), cudaMemcpyDeviceToHost, stream_));
    SAGE_CUDA(cudaStreamSynchronize(stream_));
    *occupied = icp_pin_.p->stat_occupied;
    *candidates = icp_pin_.p->stat_candidates;
}

// development aid: per-block globaltimer stamps of the LAST profiled iteration (4 per block + 4 trailing)
size_t VoxelMapGPU::debug_timeline(unsigned long long *out, size_t cap) {
    set_device();
    dbg_on_ = true;
    dbg_.ensure((size_t)kDbg * 4096 + 8);
    SAGE_CUDA(cudaStreamSynchronize(stream_));
    const int g = nn_grid_ > tile_grid_ ? nn_grid_ : tile_grid_;
    const size_t n = (size_t)kDbg * (g > 0 ? g : 1) + 8;
    if (out && cap >= n) SAGE_CUDA(cudaMemcpy(out, dbg_.p, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    return n;
}

void VoxelMapGPU::search_work(const double *xyzl, size_t n, double max_dist, double sem_th, unsigned long long *scanned,
                              unsigned long long *probes, unsigned long long *exact, unsigned long long *heavy, unsigned long long *staged) {
    set_device();
    *scanned = *probes = *exact = 0;
    if (heavy) *heavy = 0;
    if (staged) *staged = 0;
    if (n == 0 || empty()) return;
    double4 *d = stage_points(xyzl, n);
    icp_.ensure(1);
    icp_pin_.ensure(1);
    SAGE_LAUNCH(icp_init_kernel, 1, 1, 0, stream_, icp_.p, pose_identity(), 1, 0.0);
    launch_iteration(d, n, max_dist, 1.0, sem_th, 2, nullptr, nullptr);
    SAGE_CUDA(cudaMemcpyAsync(icp_pin_.p, icp_.p, sizeof(IcpState), cudaMemcpyDeviceToHost, stream_));
    SAGE_CUDA(cudaStreamSynchronize(stream_));
    *scanned = icp_pin_.p->stat_scanned, *probes = icp_pin_.p->stat_probes, *exact = icp_pin_.p->stat_exact;
    if (heavy) *heavy = icp_pin_.p->stat_heavy;
    if (staged) *staged = icp_pin_.p->stat_staged;
}

}  // namespace sage
