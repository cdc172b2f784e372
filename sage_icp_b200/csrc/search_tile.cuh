// Tile search: the correspondence search of large scans, with the voxel buckets staged through TMA bulk copies (sm_100a).
// Included by registration.cu (same translation unit: it shares the ranking helpers, the exact f64 path and reduce_and_step).
//
// Replaces, like nn_search_kernel, TransformPoints + VoxelHashMap::GetCorrespondences + AlignClouds (core/Registration.cpp:59-111,
// core/VoxelHashMap.cpp:48-130) — same per-query result (the f64 arg-min of the reference's 27-voxel scan), different schedule:
//
//   * Once per registration the queries are sorted by voxel inside 2x2x2-voxel cells (positions under the initial guess,
//     tile_sort.cu) and cut into UNITS: runs of at most kTileThreads queries of one cell.  A 120 k-point scan falls into a few
//     thousand cells, so the 27-neighbourhoods of a unit's queries overlap almost completely.
//   * Per Gauss-Newton iteration blocks take units from an atomic counter.  For a unit the block transforms its queries (in place,
//     as the reference does), takes the bounding box of their CURRENT home voxels grown by one voxel — the region, at most
//     kTileSlots voxels — probes the hash table once per region voxel (one thread per voxel, all probes in flight together) and
//     pulls every occupied bucket of the region into shared memory with one `cp.async.bulk` (TMA bulk copy, completion on an
//     mbarrier) per bucket: <= 640 contiguous bytes of 16-byte search records.  One round trip to L2/HBM per unit instead of one
//     per (query, voxel, 4 records).
//   * Ranking runs against shared memory in three rounds (same pruning rule, f32 error band and acceptance as nn_search_kernel,
//     DESIGN.md §4): (0) every query's home bucket, by its own thread — neighbouring lanes share the home voxel, so the loop is
//     converged and the loads broadcast; (1) its nearest still-open occupied neighbour, by its own thread; (2) whatever is still
//     open after that — few (query, bucket) pairs, very unevenly spread (a query in empty space has a dozen) — goes on a
//     block-wide pair list that the 128 threads share evenly, and every query merges the (min1, min2, arg) of its pairs.
//     On the bench scan this ranks 5 % more records than the fully sequential nearest-first order and keeps the lanes busy.
//   * The 17 sums are published per unit, added per group of kTileGroup units by the block that finishes the group's last unit,
//     and the block that finishes the last group adds the groups and takes the Gauss-Newton step: every sum is formed in an
//     order fixed by the unit list, so results are reproducible although units are scheduled dynamically.
//   * Correctness never depends on the sort: the region is computed from the queries' actual keys every iteration.  A unit whose
//     region would exceed kTileSlots voxels (queries that drifted apart, aliasing cells) is searched from global memory by
//     search_query_warp; buckets that do not fit the staging area are scanned from global memory in place.
#pragma once

namespace sage {

constexpr int kTileThreads = 128;
constexpr int kTileWarps = kTileThreads / 32;
constexpr int kTileSlots = 2 * kTileThreads;  // region voxels per unit (two table probes per thread)
constexpr int kTileCols = kTileThreads / 4;   // columns of a unit's sums (four threads share one, through two shuffles)
#ifndef SAGE_TILE_PAIRS
#define SAGE_TILE_PAIRS 256
#endif
#ifndef SAGE_TILE_GROUP
#define SAGE_TILE_GROUP 16
#endif
constexpr int kTilePairs = SAGE_TILE_PAIRS;   // (query, bucket) pairs per pooled round
constexpr int kTileGroup = SAGE_TILE_GROUP;   // units per group of the two-level, fixed-order sum (tile_sort.cu sizes the counters for >= 16)
static_assert(kTileGroup >= 16, "tile_prepare sizes the group counters for groups of at least 16 units");
constexpr uint32_t kNotStaged = 0xffffffffu;
static_assert(kTileCols == 32, "a unit's sums are reduced by one warp per sum");

// ---- mbarrier / bulk-copy PTX (sm_90+; SASS: SYNCS.*, UBLKCP) ----------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    // try_wait suspends the thread for a hardware-defined time slice; a barrier that never completes (a byte count that does not
    // match its copies: a bug, not a run-time condition) traps after ~2^22 slices instead of hanging the device
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t"
            ".reg .pred P1;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, P1;\n\t"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1u << 22)) __trap();
    }
}
// global -> shared bulk copy (bytes: multiple of 16, both addresses 16-byte aligned); completes `bytes` on the mbarrier
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}

// one thread ranks one staged bucket (four independent 16-byte shared loads in flight)
__device__ __forceinline__ void scan_bucket_smem(const float4 *s, uint32_t gbase, uint32_t cnt, float rx, float ry, float rz, float qlf,
                                                 float th32, float &min1, float &min2, uint32_t &idx1, bool &odd) {
    uint32_t j = 0;
    for (; j + 4 <= cnt; j += 4) {
        float4 h[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) h[u] = s[j + u];
#pragma unroll
        for (int u = 0; u < 4; ++u) rank_record(h[u], gbase + j + u, rx, ry, rz, qlf, th32, min1, min2, idx1, odd);
    }
    for (; j < cnt; ++j) rank_record(s[j], gbase + j, rx, ry, rz, qlf, th32, min1, min2, idx1, odd);
}

// shared state of one block of the tile kernel (static shared memory; the staging area is the dynamic part)
struct TileShared {
    double acc[kSums][kTileCols + 1];  // the current unit's sums, one column per four threads (padded: the publishing lanes read rows)
    Pose est;
    double norm;
    uint32_t gbase[kTileSlots], soff[kTileSlots], cnt[kTileSlots];  // region table: first record (global index), staging offset, records
    float4 qs[kTileThreads];                                        // the unit's queries: f32 offset from the home voxel's origin + label
    int qhs[kTileThreads];                                          // ... and the home voxel's slot in the region table
    uint32_t pair[kTilePairs];                                      // pooled round: (query << 8) | neighbour (reference enumeration index)
    float pr1[kTilePairs], pr2[kTilePairs];                         // ... and what ranking that bucket gave: best, second best (NaN:
    uint32_t pri[kTilePairs];                                       //     a record the f32 ranking cannot serve), arg of the best
    int wbox[kTileWarps][6];
    uint32_t wsum[kTileWarps];
    uint32_t unit;
    int flag;
    unsigned long long mbar;
    unsigned long long dbg_t[kDbg];  // development timeline (tools/tile_probe.py): time thread 0 spent in each phase, summed over units
};

// block-wide exclusive prefix sum of one value per thread (two barriers); `total` = the block's sum
__device__ __forceinline__ uint32_t tile_block_scan(TileShared &sh, uint32_t mine, uint32_t &total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();  // wsum of the previous scan has been read
    if (lane == 31) sh.wsum[warp] = incl;
    __syncthreads();
    uint32_t base = incl - mine;
    total = 0;
#pragma unroll
    for (int w = 0; w < kTileWarps; ++w) {
        base += w < warp ? sh.wsum[w] : 0u;
        total += sh.wsum[w];
    }
    return base;
}

// Sum k of a group: its units' partials added in unit order (the order is what makes the result reproducible).  The loads do not
// depend on each other, so all of them are issued before the first addition: one trip to L2 instead of one per four units — the
// last group of an iteration is summed on the iteration's critical path (measured: every unit more per group cost 0.28 us per
// iteration, profiles/r02w_small_scans.md §9).
__device__ __forceinline__ double tile_group_sum(const double *unit_part, uint32_t gfirst, uint32_t gsize, int k) {
    double t[kTileGroup];
#pragma unroll
    for (int i = 0; i < kTileGroup; ++i) t[i] = (uint32_t)i < gsize ? __ldcg(&unit_part[(size_t)(gfirst + i) * kSums + k]) : 0.0;
    double v = 0;
#pragma unroll
    for (int i = 0; i < kTileGroup; ++i)
        if ((uint32_t)i < gsize) v += t[i];
    return v;
}

// true (grid-uniformly) if this query set is left to the per-query kernel; sets the flag the host reads
__device__ __forceinline__ bool tile_declines(const IterParams &p, uint32_t n_units) {
    const unsigned long long rounds = (n_units + gridDim.x - 1) / gridDim.x;
    const bool thin = p.tile_fill != 0 && rounds * 15000ull + 12000ull > 30000ull + (unsigned long long)p.n * 48ull / 100ull;
    if (thin || __ldcg(&p.st->declined)) {
        if (blockIdx.x == 0 && threadIdx.x == 0) p.st->declined = 1, p.st->done = 1;
        return true;
    }
    return false;
}

// One Gauss-Newton iteration over the unit list.  `apply_est`: transform the queries by st->est first (every iteration but the
// first: tile_sort.cu has applied the initial guess while sorting).  `phase`: parity of the block's mbarrier, carried across
// iterations by the persistent kernel.
// `ls` (persistent kernel, single rank): the loop state kept on chip — the estimate comes from there, group sums go to the partial
// buffer of the iteration's parity, and nobody is elected to finish: every block takes the step after the caller's grid barrier.
template <bool COUNT>
__device__ __forceinline__ void nn_tile_iteration(const IterParams &p, TileShared &sh, float4 *stage, bool apply_est, uint32_t &phase,
                                                  unsigned long long tag, const LoopState *ls = nullptr) {
    const unsigned FULL = 0xffffffffu;
    const float INF = __int_as_float(0x7f800000);
    IcpState *st = p.st;
    if (ls == nullptr && p.respect_done && __ldcg(&st->done)) return;
    if (threadIdx.x == 0) sh.est = ls ? ls->est : load_pose_cg(&st->est);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double vs = p.voxel_size;
    const float vs32 = p.vs32, th32 = p.th32;
    const uint32_t bar = smem_addr(&sh.mbar);
    const uint32_t n_units = *p.tile_n_units;
    // The tile search pays per unit (~15 us of dependent latency for each round of gridDim.x units, ~12 us per iteration for the
    // reduction and the step); the per-query kernel pays ~30 us + 0.48 ns per query (both measured on B200, profiles/
    // r02_tile_kernel.md).  A query set that is spread so thinly over the map that its units would take longer than that — BASELINE
    // configs[4]'s uniform queries: one unit per query — is declined here, grid-uniformly and before anything is touched: the host
    // sees the flag at its next read-back and runs the per-query kernel on the (already sorted) array.  Deciding on the device keeps a
    // host round trip out of every ordinary registration.  tile_fill == 0: never decline (tests force a schedule).
    if (ls == nullptr && tile_declines(p, n_units)) return;  // (the persistent kernel asks once, before its loop)
    const uint32_t n_groups = (n_units + kTileGroup - 1) / kTileGroup;
    // Units are handed out by a counter that only grows during a registration (p.tile_ctl[0], alone on its cache line: every
    // block hits it two or three times per iteration); p.tile_ctl[32] is its value at the start of this iteration — every block
    // makes exactly one failing fetch per iteration, so the block that finishes the iteration advances it by n_units + gridDim.x.
    // (A block that is scheduled so late that the base has already advanced wraps around and leaves without a unit.)
    const uint32_t fetch_base = ls ? (uint32_t)ls->it * (n_units + gridDim.x) : __ldcg(p.tile_ctl + 32);
    double *group_partials = p.partials + (ls ? (size_t)(ls->it & 1) * kSums * n_groups : 0);
    // Hand-out order: largest units first (tile_sort.cu), so that what a block picks up last is small.  The order only schedules:
    // every unit's sums go to the unit's own slot.
    auto fetch_unit = [&]() -> uint32_t {  // thread 0 only
        const uint32_t t = atomicAdd(p.tile_ctl, 1u) - fetch_base;
        return (t < n_units && p.tile_order != nullptr) ? p.tile_order[t] : t;
    };
    // development timeline: thread 0 adds the time since its previous stamp to phase k (the block barriers align the warps)
    unsigned long long t_last = 0;
    if (p.dbg && threadIdx.x == 0) {
        for (int k = 0; k < kDbg; ++k) sh.dbg_t[k] = 0;
        t_last = gtime();
        sh.dbg_t[0] = t_last;
    }
#define TILE_STAMP(k)                              \
    do {                                           \
        if (p.dbg && threadIdx.x == 0) {           \
            const unsigned long long t_ = gtime(); \
            sh.dbg_t[k] += t_ - t_last;            \
            t_last = t_;                           \
        }                                          \
    } while (0)
    unsigned long long n_ranked = 0, n_probes = 0, n_exact = 0, n_pooled = 0, n_staged = 0;  // work counters (COUNT launches only)

    uint32_t next_unit = 0;  // thread 0: the unit after the current one, drawn while the current one is being finished
    if (threadIdx.x == 0) sh.unit = fetch_unit();
    __syncthreads();
    for (;;) {
        const uint32_t u = sh.unit;  // set before the last barrier
        if (u >= n_units) break;
        TILE_STAMP(8);
        const uint32_t ubeg = p.tile_units[u], uend = p.tile_units[u + 1];  // at most kTileThreads queries by construction
        const uint32_t q = ubeg + threadIdx.x;
        const bool valid = q < uend;
        // ---- A: transform (in place, core/Registration.cpp:133), home voxel, f32 query ---------------------------------------
        int kx, ky, kz;
        float bx, by, bz, qlf;
        {
            double sx = 0, sy = 0, sz = 0, sl = 0;
            if (valid) {
                const double4 s = ld256(p.src + q);
                sx = s.x, sy = s.y, sz = s.z, sl = s.w;
                if (apply_est) {
                    const Pose est = sh.est;
                    pose_act(est, s.x, s.y, s.z, sx, sy, sz);
                    st256(p.src + q, make_double4(sx, sy, sz, sl));
                }
            }
            kx = trunc_div(sx, vs), ky = trunc_div(sy, vs), kz = trunc_div(sz, vs);
            bx = hot_offset(sx, kx, vs), by = hot_offset(sy, ky, vs), bz = hot_offset(sz, kz, vs);
            qlf = hot_label(sl);
        }
        // `odd`: queries the f32 ranking cannot serve (they take the exact f64 path and stay out of the region)
        bool odd = !p.fast_ok || (qlf != qlf) || !(fabsf(bx) <= 2.0f * vs32 && fabsf(by) <= 2.0f * vs32 && fabsf(bz) <= 2.0f * vs32) ||
                   !key_in_range(kx, ky, kz);
        const bool fast = valid && !odd;
        // ---- B: the region = bounding box of the home voxels, grown by one voxel -----------------------------------------------
        {
            const int big = 0x7fffffff;
            const int mnx = __reduce_min_sync(FULL, fast ? kx : big), mny = __reduce_min_sync(FULL, fast ? ky : big),
                      mnz = __reduce_min_sync(FULL, fast ? kz : big);
            const int mxx = __reduce_max_sync(FULL, fast ? kx : -big), mxy = __reduce_max_sync(FULL, fast ? ky : -big),
                      mxz = __reduce_max_sync(FULL, fast ? kz : -big);
            if (lane == 0) {
                sh.wbox[warp][0] = mnx, sh.wbox[warp][1] = mny, sh.wbox[warp][2] = mnz;
                sh.wbox[warp][3] = mxx, sh.wbox[warp][4] = mxy, sh.wbox[warp][5] = mxz;
            }
        }
        __syncthreads();  // boxes of all warps; the transformed points of this unit are visible to the block
        TILE_STAMP(1);
        int lox = sh.wbox[0][0], loy = sh.wbox[0][1], loz = sh.wbox[0][2], hix = sh.wbox[0][3], hiy = sh.wbox[0][4], hiz = sh.wbox[0][5];
#pragma unroll
        for (int w = 1; w < kTileWarps; ++w) {
            lox = min(lox, sh.wbox[w][0]), loy = min(loy, sh.wbox[w][1]), loz = min(loz, sh.wbox[w][2]);
            hix = max(hix, sh.wbox[w][3]), hiy = max(hiy, sh.wbox[w][4]), hiz = max(hiz, sh.wbox[w][5]);
        }
        const bool any_fast = lox <= hix;  // block-uniform
        lox -= 1, loy -= 1, loz -= 1;
        // extents (keys are within +-2^20, so these fit easily); tiled: the region fits the table, else global fallback
        const long long ex = (long long)hix - lox + 2, ey = (long long)hiy - loy + 2, ez = (long long)hiz - loz + 2;
        const bool tiled = any_fast && ex * ey * ez <= (long long)kTileSlots;
        const int dz = tiled ? (int)ez : 1, dyz = tiled ? (int)(ey * ez) : 1, vol = tiled ? (int)(ex * ey * ez) : 0;
        // ---- C: one table probe per region voxel, staging offsets by a block scan, one bulk copy per occupied bucket -------------
        {
            uint32_t c[2] = {0, 0}, g[2] = {0, 0};
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int slot = threadIdx.x + e * kTileThreads;
                if (slot < vol) {
                    const int iz = slot % dz, iy = (slot / dz) % (dyz / dz), ix = slot / dyz;
                    const int nx = lox + ix, ny = loy + iy, nz = loz + iz;
                    uint32_t blk = 0, cn = 0;
                    if (key_in_range(nx, ny, nz) && tbl_find(p.tbl, p.mask, pack_key(nx, ny, nz), blk, cn) && cn > 0)
                        c[e] = cn, g[e] = blk * (uint32_t)p.stride;
                    if (COUNT) n_probes += 1;
                }
            }
            uint32_t total;
            const uint32_t base = tile_block_scan(sh, c[0] + c[1], total);
            uint32_t off[2] = {base, base + c[0]}, bytes = 0;
            bool staged[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                staged[e] = c[e] > 0 && off[e] + c[e] <= p.tile_stage_cap;
                const int slot = threadIdx.x + e * kTileThreads;
                sh.gbase[slot] = g[e], sh.cnt[slot] = c[e], sh.soff[slot] = staged[e] ? off[e] : kNotStaged;
                if (staged[e]) bytes += c[e] * 16u;
            }
            if (COUNT) n_staged += bytes / 16u;
            if (bytes)
                mbar_arrive_expect_tx(bar, bytes);
            else
                mbar_arrive(bar);
#pragma unroll
            for (int e = 0; e < 2; ++e)
                if (staged[e]) bulk_g2s(smem_addr(stage + off[e]), p.blk_hot + g[e], c[e] * 16u, bar);
        }
        TILE_STAMP(2);
        __syncthreads();        // region table complete
        mbar_wait(bar, phase);  // every staged bucket has landed
        phase ^= 1u;
        TILE_STAMP(3);

        // ---- D: ranking against the staged region --------------------------------------------------------------------------------
        uint32_t widx = kNil;
        float min1 = INF, min2 = INF;
        uint32_t idx1 = kNil, visited = 1u << 13, open = 0;
        const bool ranked = fast && tiled;
        const int hs = ranked ? ((kx - lox) * (dyz / dz) + (ky - loy)) * dz + (kz - loz) : 0;
        auto scan_slot = [&](int slot, float rx, float ry, float rz, float ql, float &m1, float &m2, uint32_t &i1, bool &od) {
            const uint32_t cn = sh.cnt[slot], o = sh.soff[slot], g = sh.gbase[slot];
            if (COUNT) n_ranked += cn;
            if (o != kNotStaged)
                scan_bucket_smem(stage + o, g, cn, rx, ry, rz, ql, th32, m1, m2, i1, od);
            else
                scan_voxel_thread(p.blk_hot, g, cn, rx, ry, rz, ql, th32, m1, m2, i1, od);
        };
        if (ranked) {
            // round 0: the home bucket
            if (sh.cnt[hs]) scan_slot(hs, bx, by, bz, qlf, min1, min2, idx1, odd);
            float sxm, sxp, sym, syp, szm, szp;
            axis_bounds(bx, kx, vs32, p.box_margin, p.smin32, sxm, sxp);
            axis_bounds(by, ky, vs32, p.box_margin, p.smin32, sym, syp);
            axis_bounds(bz, kz, vs32, p.box_margin, p.smin32, szm, szp);
            // round 1: the nearest occupied neighbour whose box can still beat (or tie) the best so far
            float bound = prune_bound(p, min1);
            {
                float best_lb = INF;
                int nn = -1;
#pragma unroll
                for (int ox = 0; ox < 3; ++ox) {
                    const float ax = ox == 0 ? sxm : (ox == 2 ? sxp : 0.0f);
#pragma unroll
                    for (int oy = 0; oy < 3; ++oy) {
                        const float ay = ax + (oy == 0 ? sym : (oy == 2 ? syp : 0.0f));
#pragma unroll
                        for (int oz = 0; oz < 3; ++oz) {
                            const float az = ay + (oz == 0 ? szm : (oz == 2 ? szp : 0.0f));
                            const int id = ox * 9 + oy * 3 + oz;
                            if (id != 13 && az <= bound && sh.cnt[hs + (ox - 1) * dyz + (oy - 1) * dz + (oz - 1)] != 0) {
                                open |= 1u << id;
                                if (az < best_lb) best_lb = az, nn = id;
                            }
                        }
                    }
                }
                if (nn >= 0) {
                    visited |= 1u << nn;
                    const int ox = nn / 9 - 1, oy = (nn / 3) % 3 - 1, oz = nn % 3 - 1;
                    scan_slot(hs + ox * dyz + oy * dz + oz, bx - (float)ox * vs32, by - (float)oy * vs32, bz - (float)oz * vs32, qlf, min1, min2,
                              idx1, odd);
                    bound = prune_bound(p, min1);
                }
            }
            // round 2 candidates: what was open before round 1, is not visited, and survives the tightened bound
            open &= ~visited;
            if (odd) open = 0;  // met a record the f32 ranking cannot serve: exact path below
            if (open) {
                uint32_t still = 0;
#pragma unroll
                for (int ox = 0; ox < 3; ++ox) {
                    const float ax = ox == 0 ? sxm : (ox == 2 ? sxp : 0.0f);
#pragma unroll
                    for (int oy = 0; oy < 3; ++oy) {
                        const float ay = ax + (oy == 0 ? sym : (oy == 2 ? syp : 0.0f));
#pragma unroll
                        for (int oz = 0; oz < 3; ++oz) {
                            const float az = ay + (oz == 0 ? szm : (oz == 2 ? szp : 0.0f));
                            if (az <= bound) still |= 1u << (ox * 9 + oy * 3 + oz);
                        }
                    }
                }
                open &= still;
            }
            sh.qs[threadIdx.x] = make_float4(bx, by, bz, qlf);
            sh.qhs[threadIdx.x] = hs;
        }
        TILE_STAMP(4);
        // ---- E: round 2, pooled: the block's open (query, bucket) pairs, dealt evenly over its threads ---------------------------
        {
            const uint32_t mine = (uint32_t)__popc(open);
            uint32_t total;
            const uint32_t base = tile_block_scan(sh, mine, total);  // block-uniform total
            if (COUNT) n_pooled += mine;
            for (uint32_t lo = 0; lo < total; lo += kTilePairs) {
                const uint32_t hi = min(lo + (uint32_t)kTilePairs, total);
                // emit: the k-th open neighbour of this query is pair base + k
                {
                    uint32_t m = open, k = base;
                    while (m) {
                        const int id = __ffs(m) - 1;
                        m &= m - 1;
                        if (k >= lo && k < hi) sh.pair[k - lo] = ((uint32_t)threadIdx.x << 8) | (uint32_t)id;
                        ++k;
                    }
                }
                __syncthreads();
                for (uint32_t i = lo + threadIdx.x; i < hi; i += kTileThreads) {
                    const uint32_t e = sh.pair[i - lo];
                    const int qi = (int)(e >> 8), id = (int)(e & 0xffu);
                    const float4 qv = sh.qs[qi];
                    const int ox = id / 9 - 1, oy = (id / 3) % 3 - 1, oz = id % 3 - 1;
                    float r1 = INF, r2 = INF;
                    uint32_t ri = kNil;
                    bool rodd = false;
                    scan_slot(sh.qhs[qi] + ox * dyz + oy * dz + oz, qv.x - (float)ox * vs32, qv.y - (float)oy * vs32, qv.z - (float)oz * vs32, qv.w,
                              r1, r2, ri, rodd);
                    sh.pr1[i - lo] = r1, sh.pr2[i - lo] = rodd ? __int_as_float(0x7fc00000) : r2, sh.pri[i - lo] = ri;
                }
                __syncthreads();
                // merge: (min1, min2, arg) of the union of two record sets; a tie keeps min2 == min1, which the band test rejects
                {
                    uint32_t k = base;
                    for (uint32_t m = open; m; m &= m - 1, ++k) {
                        if (k < lo || k >= hi) continue;
                        const float a1 = sh.pr1[k - lo], a2 = sh.pr2[k - lo];
                        if (a2 != a2) odd = true;
                        if (a1 < min1) {
                            min2 = fminf(min1, a2), min1 = a1, idx1 = sh.pri[k - lo];
                        } else {
                            min2 = fminf(min2, a1);
                        }
                    }
                }
                if (hi < total) __syncthreads();  // the pair arrays are rewritten by the next round
            }
        }
        // ---- decide: unique within the error band => idx1 IS the f64 arg-min ------------------------------------------------------
        if (ranked && !odd && min1 < INF) {
            if (min2 > band_limit(p, min1))
                widx = idx1;
            else
                odd = true;
        }
        TILE_STAMP(5);
        // ---- F: a unit whose region does not fit (never on a sorted scan): every query through the global-memory warp search ------
        if (!tiled) {
            unsigned fm = __ballot_sync(FULL, fast);
            unsigned long long d0 = 0, d1 = 0, d2 = 0;
            while (fm) {
                const int l = __ffs(fm) - 1;
                fm &= fm - 1;
                const double4 c = ld256(p.src + ubeg + warp * 32 + l);
                const uint32_t w = search_query_warp<false>(p, lane, c, Carried{INF, INF, kNil, 0u}, d0, d1, d2);
                if (lane == l) widx = w;
            }
        }
        // ---- G: queries the f32 ranking could not decide: the reference's own f64 sequence over all 27 voxels ----------------------
        {
            unsigned em = __ballot_sync(FULL, valid && odd);
            while (em) {
                const int l = __ffs(em) - 1;
                em &= em - 1;
                const double4 c = ld256(p.src + ubeg + warp * 32 + l);
                const uint32_t w = nn_exact<32>(p, FULL, lane, c.x, c.y, c.z, c.w, trunc_div(c.x, vs), trunc_div(c.y, vs), trunc_div(c.z, vs));
                if (lane == l) widx = w;
                if (COUNT && lane == 0) n_exact += 1;
            }
        }
        TILE_STAMP(6);
        // The next unit is drawn HERE, not at the top of the loop: a draw is a claim, and claims made at the start of a unit go to the
        // blocks that started first — which hold the largest units (hand-out order) — instead of the blocks that finish first.  With
        // fewer units than twice the grid (a 15 000-query shard of an 8-GPU run: ~700 units on 592 blocks) the eager draw gave the
        // blocks with a full unit a second one while two thirds of the grid idled (31 us to the last unit where one full unit takes
        // 17: profiles/r02w_small_scans.md §6).  Phase H's record fetch still hides the atomic's round trip.
        if (threadIdx.x == 0) next_unit = fetch_unit();
        // ---- H: acceptance on the exact records, residual, weight, the 16 sums + pair count -----------------------------------------
        {
            double a[kSums];
#pragma unroll
            for (int k = 0; k < kSums; ++k) a[k] = 0.0;
            const bool any_pair = __any_sync(FULL, widx != kNil);
            double4 nb = make_double4(0, 0, 0, 0);
            bool ok = false;
            if (widx != kNil) {
                const double4 s = ld256(p.src + q);
                nb = ldg256(p.blk_pts + widx);
                ok = accept_pair(nb, s.x, s.y, s.z, p.max_dist);
                if (ok) {
                    const double rx = s.x - nb.x, ry = s.y - nb.y, rz = s.z - nb.z;
                    const double r2 = (rx * rx + ry * ry) + rz * rz;
                    const double den = p.kern + r2;
                    const double w = (p.kern * p.kern) / (den * den);
                    const double wx = w * s.x, wy = w * s.y, wz = w * s.z;
                    a[0] = w, a[1] = wx, a[2] = wy, a[3] = wz;
                    a[4] = wx * s.x, a[5] = wy * s.y, a[6] = wz * s.z, a[7] = wx * s.y, a[8] = wx * s.z, a[9] = wy * s.z;
                    a[10] = w * rx, a[11] = w * ry, a[12] = w * rz;
                    a[13] = w * (s.y * rz - s.z * ry), a[14] = w * (s.z * rx - s.x * rz), a[15] = w * (s.x * ry - s.y * rx);
                    a[16] = 1.0;
                }
            }
            if (p.tgt_out != nullptr && valid) {  // correspondences in the caller's order (tile_perm: sorted position -> input index)
                const uint32_t dst = p.tile_perm[q];
                st256(p.tgt_out + dst, nb);
                p.matched_out[dst] = ok ? 1 : 0;
            }
            if (any_pair) {
#pragma unroll
                for (int k = 0; k < kSums; ++k) {
                    a[k] += __shfl_xor_sync(FULL, a[k], 1);
                    a[k] += __shfl_xor_sync(FULL, a[k], 2);
                }
            }
            if ((lane & 3) == 0) {  // every column is written for every unit (zeros from a warp without pairs)
#pragma unroll
                for (int k = 0; k < kSums; ++k) sh.acc[k][threadIdx.x >> 2] = a[k];
            }
        }
        TILE_STAMP(7);
        // ---- I: publish the unit's sums; the block that completes a group adds it; the one that completes the last group steps ------
        const uint32_t g = u / kTileGroup;
        const uint32_t gfirst = g * kTileGroup, gsize = min((uint32_t)kTileGroup, n_units - gfirst);
        if (ls != nullptr) {
            // Every block takes the step itself (persistent loop, single rank): nobody has to be elected, so warp 0 publishes the unit —
            // and, when it completes a group, adds the group — on its own while the other warps start on the next unit; it catches
            // up at that unit's first barrier (phase A takes them about as long as the publish takes warp 0).  The columns it reads are
            // rewritten only in phase H of the next unit, several block barriers later; the caller's grid barrier orders the group sums.
            if (threadIdx.x == 0) sh.unit = next_unit;
            __syncthreads();  // every warp has written its columns; the next unit is known to all
            if (warp == 0) {
                if (lane < kSums) {
                    double v = 0;
#pragma unroll 8
                    for (int c = 0; c < kTileCols; ++c) v += sh.acc[lane][c];
                    p.tile_unit_part[(size_t)u * kSums + lane] = v;
                    __threadfence();
                }
                __syncwarp();
                uint32_t last = 0;
                if (lane == 0) last = (atomicAdd(&p.tile_group_cnt[g], 1u) == gsize - 1) ? 1u : 0u;
                last = __shfl_sync(FULL, last, 0);
                if (last) {
                    __threadfence();
                    if (lane < kSums) group_partials[(size_t)lane * n_groups + g] = tile_group_sum(p.tile_unit_part, gfirst, gsize, lane);  // [sum][group]
                    if (lane == 0) p.tile_group_cnt[g] = 0;  // ready for the next iteration
                }
            }
            if (p.dbg && threadIdx.x == 0) sh.dbg_t[10] += 1, sh.dbg_t[11] += uend - ubeg;
            continue;
        }
        __syncthreads();  // every warp has written its columns
        if (warp == 0) {  // one warp publishes (lane k adds the 32 columns of sum k in order); the others go and wait at the barrier
            if (lane < kSums) {
                double v = 0;
#pragma unroll 8
                for (int c = 0; c < kTileCols; ++c) v += sh.acc[lane][c];
                p.tile_unit_part[(size_t)u * kSums + lane] = v;
                __threadfence();
            }
            __syncwarp();
            if (lane == 0) {
                sh.flag = (atomicAdd(&p.tile_group_cnt[g], 1u) == gsize - 1) ? 1 : 0;
                sh.unit = next_unit;
            }
        }
        __syncthreads();  // flag and the next unit are set; everybody is done with this unit's shared state
        if (sh.flag) {
            __threadfence();
            if (threadIdx.x < kSums) {
                group_partials[(size_t)threadIdx.x * n_groups + g] = tile_group_sum(p.tile_unit_part, gfirst, gsize, (int)threadIdx.x);  // [sum][group]
                __threadfence();
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                p.tile_group_cnt[g] = 0;  // ready for the next iteration
                sh.flag = (ls == nullptr && atomicAdd(&st->ticket, 1u) == n_groups - 1) ? 2 : 0;
            }
            __syncthreads();
            if (sh.flag == 2) {
                if (threadIdx.x == 0) p.tile_ctl[32] = fetch_base + n_units + gridDim.x;  // every unit of this iteration is done
                reduce_and_step(p, n_groups, sh.est, sh.norm, tag);
            }
            __syncthreads();  // sh.flag is rewritten by the next unit's publish
        }
        if (p.dbg && threadIdx.x == 0) sh.dbg_t[10] += 1, sh.dbg_t[11] += uend - ubeg;
    }
    if (p.dbg && threadIdx.x == 0) {
        sh.dbg_t[9] = gtime();
        for (int k = 0; k < kDbg; ++k) p.dbg[(size_t)kDbg * blockIdx.x + k] = sh.dbg_t[k];
    }
#undef TILE_STAMP
    if (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_ranked += __shfl_xor_sync(FULL, n_ranked, o), n_probes += __shfl_xor_sync(FULL, n_probes, o);
            n_exact += __shfl_xor_sync(FULL, n_exact, o), n_pooled += __shfl_xor_sync(FULL, n_pooled, o);
            n_staged += __shfl_xor_sync(FULL, n_staged, o);
        }
        if (lane == 0) {
            atomicAdd(&st->stat_scanned, n_ranked), atomicAdd(&st->stat_probes, n_probes), atomicAdd(&st->stat_exact, n_exact);
            atomicAdd(&st->stat_heavy, n_pooled), atomicAdd(&st->stat_staged, n_staged);
        }
    }
}

// One launch = one Gauss-Newton iteration (`iteration0`: the first one of a registration).
// MINB: resident blocks per SM the register allocation aims at.
template <int MINB, bool COUNT = false>
__global__ void __launch_bounds__(kTileThreads, MINB) nn_tile_kernel(IterParams p, int iteration0) {
    extern __shared__ __align__(128) float4 s_tile_stage[];
    __shared__ TileShared sh;
    if (threadIdx.x == 0) mbar_init(smem_addr(&sh.mbar), kTileThreads);
    __syncthreads();
    uint32_t phase = 0;
    nn_tile_iteration<COUNT>(p, sh, s_tile_stage, !iteration0, phase, p.xchg_tag);
}

// The whole Gauss-Newton loop in one cooperative launch (grid barrier between iterations; the fused peer exchange, when there
// are several ranks, runs inside the loop with one tag per iteration).
template <int MINB>
__global__ void __launch_bounds__(kTileThreads, MINB) nn_tile_persistent_kernel(IterParams p, int max_iterations) {
    extern __shared__ __align__(128) float4 s_tile_stage[];
    __shared__ TileShared sh;
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    if (threadIdx.x == 0) mbar_init(smem_addr(&sh.mbar), kTileThreads);
    __syncthreads();
    uint32_t phase = 0;
    if (p.step_everywhere) {  // single rank: every block takes the step, the loop state stays on chip (registration.cu)
        __shared__ LoopState ls;
        const uint32_t n_units = *p.tile_n_units;
        if (tile_declines(p, n_units)) return;
        const uint32_t n_groups = (n_units + kTileGroup - 1) / kTileGroup;
        loop_state_init(p, ls);
        while (!ls.done) {
            nn_tile_iteration<false>(p, sh, s_tile_stage, ls.it > 0, phase, 0ull, &ls);
            grid.sync();
            loop_step_everywhere(p, ls, p.partials + (size_t)(ls.it & 1) * kSums * n_groups, n_groups, max_iterations);
        }
        loop_state_commit(p, ls);
        return;
    }
    for (int i = 0; i < max_iterations; ++i) {
        nn_tile_iteration<false>(p, sh, s_tile_stage, i > 0, phase, p.xchg_tag + (unsigned long long)i);
        grid.sync();
        if (*reinterpret_cast<volatile int *>(&p.st->done)) break;
    }
}

}  // namespace sage
