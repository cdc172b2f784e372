// Tile search: the correspondence search of large scans, with the voxel buckets staged through TMA bulk copies (sm_100a).
// Included by registration.cu (same translation unit: it shares the ranking helpers, the exact f64 path and finish_iteration).
//
// Replaces, like nn_search_kernel, TransformPoints + VoxelHashMap::GetCorrespondences + AlignClouds (core/Registration.cpp:59-111,
// core/VoxelHashMap.cpp:48-130) — same per-query result (the f64 arg-min of the reference's 27-voxel scan), different schedule:
//
//   * Once per registration the queries are sorted by the 2x2x2-voxel cell of their position under the initial guess
//     (tile_sort.cu) and cut into UNITS: runs of at most kTileThreads queries of one cell.  A 120 k-point scan falls into a few
//     thousand cells, so the 27-neighbourhoods of a unit's queries overlap almost completely.
//   * Per Gauss-Newton iteration a block takes units round-robin.  For each it transforms the unit's queries (in place, as the
//     reference does), takes the bounding box of their CURRENT home voxels grown by one voxel — the region, at most kTileSlots
//     voxels — probes the hash table once per region voxel (one thread per voxel, all probes in flight together), and pulls
//     every occupied bucket of the region into shared memory with one `cp.async.bulk` (TMA bulk copy, completion on an mbarrier)
//     per bucket: <= 640 contiguous bytes of 16-byte search records each.  One round trip to L2/HBM per unit instead of one
//     per (query, voxel, 4 records).
//   * Each query is then ranked by its own thread against shared memory: home voxel, then neighbours nearest bounding box first
//     while one can still beat the best so far (same pruning rule, same f32 error band, same acceptance as nn_search_kernel;
//     DESIGN.md §4).  Queries that still have open neighbours after `light_probes` visits are finished by their whole warp
//     (lane = record), so a far query with 20 voxels to look at does not hold its warp for 20 serial scans.
//   * Correctness never depends on the sort: the region is computed from the queries' actual keys every iteration.  A unit whose
//     region would exceed kTileSlots voxels (queries that drifted apart, aliasing cells) is searched from global memory by
//     search_query_warp; buckets that do not fit the staging area are scanned from global memory in place.
#pragma once

namespace sage {

constexpr int kTileThreads = 128;
constexpr int kTileWarps = kTileThreads / 32;
constexpr int kTileSlots = 2 * kTileThreads;  // region voxels per unit (two table probes per thread)
constexpr int kTileCols = kTileThreads / 4;   // columns of running sums (four threads share one, through two shuffles)
constexpr uint32_t kNotStaged = 0xffffffffu;
static_assert(kTileCols % 32 == 0, "finish_iteration reduces whole warps of columns");

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
    double acc[kSums][kTileCols];
    Pose est;
    double norm;
    int last;
    uint32_t gbase[kTileSlots], soff[kTileSlots], cnt[kTileSlots];  // region table: first record (global index), staging offset, records
    int wbox[kTileWarps][6];
    uint32_t wsum[kTileWarps];
    unsigned long long mbar;
    unsigned long long dbg_t[kDbg];  // development timeline (tools/tile_probe.py): time thread 0 spent in each phase, summed over units
};

// Warp-cooperative end of one query's search inside a staged region: the neighbours the thread phase left open, nearest box
// first, each bucket ranked by the whole warp (lane = record), the prune bound re-tightened after every bucket.  All lanes pass
// the same arguments; returns (identically on every lane) the winner's record index, kNil for "no candidate", and sets
// `ambiguous` when the f32 ranking cannot decide (the caller re-ranks in f64).
__device__ __noinline__ uint32_t tile_finish_query_warp(const IterParams &p, const TileShared &sh, const float4 *stage, int lane, int hs,
                                                        int dyz, int dz, float bx, float by, float bz, int kx, int ky, int kz, float qlf,
                                                        Carried cs, bool &ambiguous) {
    const unsigned FULL = 0xffffffffu;
    const float INF = __int_as_float(0x7f800000);
    const float vs32 = p.vs32, th32 = p.th32;
    // lane = neighbour voxel in the reference's enumeration order
    const int ox = lane / 9 - 1, oy = (lane / 3) % 3 - 1, oz = lane % 3 - 1;
    float lb = INF;
    uint32_t vcnt = 0, vg = 0, vo = kNotStaged;
    if (lane < 27 && !((cs.visited >> lane) & 1u)) {
        const int slot = hs + ox * dyz + oy * dz + oz;
        vcnt = sh.cnt[slot];
        if (vcnt) {
            float sxm, sxp, sym, syp, szm, szp;
            axis_bounds(bx, kx, vs32, p.box_margin, p.smin32, sxm, sxp);
            axis_bounds(by, ky, vs32, p.box_margin, p.smin32, sym, syp);
            axis_bounds(bz, kz, vs32, p.box_margin, p.smin32, szm, szp);
            lb = (ox < 0 ? sxm : (ox > 0 ? sxp : 0.0f)) + (oy < 0 ? sym : (oy > 0 ? syp : 0.0f)) + (oz < 0 ? szm : (oz > 0 ? szp : 0.0f));
            vg = sh.gbase[slot], vo = sh.soff[slot];
        }
    }
    // lane 0 carries the thread phase's result into the reduction
    float min1 = lane == 0 ? cs.min1 : INF, min2 = lane == 0 ? cs.min2 : INF;
    uint32_t idx1 = lane == 0 ? cs.idx1 : kNil;
    float gbest = cs.min1;
    bool odd = false;
    while (true) {
        // nearest open box (lb >= 0, so the unsigned order of the bits is the float order; INF = closed)
        const unsigned nb = __reduce_min_sync(FULL, __float_as_uint(lb));
        const float nlb = __uint_as_float(nb);
        if (nb == 0x7f800000u || !(nlb <= prune_bound(p, gbest))) break;  // nothing open, or nothing open can still win
        const int l = __ffs(__ballot_sync(FULL, __float_as_uint(lb) == nb)) - 1;
        const uint32_t c = __shfl_sync(FULL, vcnt, l), g = __shfl_sync(FULL, vg, l), o = __shfl_sync(FULL, vo, l);
        const float rx = bx - (float)(l / 9 - 1) * vs32, ry = by - (float)((l / 3) % 3 - 1) * vs32, rz = bz - (float)(l % 3 - 1) * vs32;
        if (lane == l) lb = INF;
        for (uint32_t j = lane; j < c; j += 32) {
            const float4 h = (o != kNotStaged) ? stage[o + j] : __ldg(p.blk_hot + g + j);
            rank_record(h, g + j, rx, ry, rz, qlf, th32, min1, min2, idx1, odd);
        }
        gbest = __uint_as_float(__reduce_min_sync(FULL, __float_as_uint(min1)));
    }
    ambiguous = __any_sync(FULL, odd);
    uint32_t widx = kNil;
    if (!ambiguous && gbest < INF) {
        const float T = band_limit(p, gbest);
        const unsigned in1 = __ballot_sync(FULL, min1 <= T), in2 = __ballot_sync(FULL, min2 <= T);
        if (__popc(in1) == 1 && in2 == 0)
            widx = __shfl_sync(FULL, idx1, __ffs(in1) - 1);
        else
            ambiguous = true;
    }
    return widx;
}

// One Gauss-Newton iteration over the unit list.  `apply_est`: transform the queries by st->est first (every iteration but the
// first: tile_sort.cu has applied the initial guess while sorting).  `phase`: parity of the block's mbarrier, carried across
// iterations by the persistent kernel.
template <bool COUNT>
__device__ __forceinline__ void nn_tile_iteration(const IterParams &p, TileShared &sh, float4 *stage, bool apply_est, uint32_t &phase,
                                                  unsigned long long tag) {
    const unsigned FULL = 0xffffffffu;
    const float INF = __int_as_float(0x7f800000);
    IcpState *st = p.st;
    if (p.respect_done && __ldcg(&st->done)) return;
    if (threadIdx.x == 0) sh.est = load_pose_cg(&st->est);
    for (int k = threadIdx.x; k < kSums * kTileCols; k += kTileThreads) (&sh.acc[0][0])[k] = 0.0;
    __syncthreads();

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double vs = p.voxel_size;
    const float vs32 = p.vs32, th32 = p.th32;
    const uint32_t bar = smem_addr(&sh.mbar);
    const uint32_t n_units = *p.tile_n_units;
    // development timeline: thread 0 adds the time since its previous stamp to phase k (the block barriers align the warps)
    unsigned long long t_last = 0;
    if (p.dbg && threadIdx.x == 0) {
        for (int k = 0; k < kDbg; ++k) sh.dbg_t[k] = 0;
        t_last = gtime();
        sh.dbg_t[0] = t_last;
    }
#define TILE_STAMP(k)                             \
    do {                                          \
        if (p.dbg && threadIdx.x == 0) {          \
            const unsigned long long t_ = gtime(); \
            sh.dbg_t[k] += t_ - t_last;           \
            t_last = t_;                          \
        }                                         \
    } while (0)
    unsigned long long n_ranked = 0, n_probes = 0, n_exact = 0, n_heavy = 0, n_staged = 0;  // work counters (COUNT launches only)
    for (uint32_t u = blockIdx.x; u < n_units; u += gridDim.x) {
        const uint32_t ubeg = p.tile_units[u], uend = p.tile_units[u + 1];
        for (uint32_t c0 = ubeg; c0 < uend; c0 += kTileThreads) {  // one pass by construction (units hold <= kTileThreads queries)
            const uint32_t q = c0 + threadIdx.x;
            const bool valid = q < uend;
            // ---- A: transform (in place, core/Registration.cpp:133), home voxel, f32 query -------------------------------------
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
            // ---- B: the region = bounding box of the home voxels, grown by one voxel ---------------------------------------------
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
            __syncthreads();  // (1) boxes of all warps; the transformed points of this unit are visible to the block
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
            // ---- C: one table probe per region voxel, staging offsets by a block scan, one bulk copy per occupied bucket -----------
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
                const uint32_t mine = c[0] + c[1];
                uint32_t incl = mine;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(FULL, incl, o);
                    if (lane >= o) incl += t;
                }
                if (lane == 31) sh.wsum[warp] = incl;
                __syncthreads();  // (2)
                uint32_t base = incl - mine;
#pragma unroll
                for (int w = 0; w < kTileWarps; ++w) base += w < warp ? sh.wsum[w] : 0u;
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
            __syncthreads();  // (3) region table complete
            mbar_wait(bar, phase);  // every staged bucket has landed
            phase ^= 1u;
            TILE_STAMP(3);

            // ---- D: thread per query against the staged region ---------------------------------------------------------------------
            uint32_t widx = kNil;
            bool heavy = false;
            float min1 = INF, min2 = INF;
            uint32_t idx1 = kNil, visited = 1u << 13;
            const int hs = tiled && fast ? ((kx - lox) * (dyz / dz) + (ky - loy)) * dz + (kz - loz) : 0;
            if (fast && tiled) {
                auto scan_slot = [&](int slot, float rx, float ry, float rz) {
                    const uint32_t cn = sh.cnt[slot];
                    if (cn == 0) return false;
                    const uint32_t o = sh.soff[slot], g = sh.gbase[slot];
                    if (COUNT) n_ranked += cn;
                    if (o != kNotStaged)
                        scan_bucket_smem(stage + o, g, cn, rx, ry, rz, qlf, th32, min1, min2, idx1, odd);
                    else
                        scan_voxel_thread(p.blk_hot, g, cn, rx, ry, rz, qlf, th32, min1, min2, idx1, odd);
                    return true;
                };
                scan_slot(hs, bx, by, bz);
                float sxm, sxp, sym, syp, szm, szp;
                axis_bounds(bx, kx, vs32, p.box_margin, p.smin32, sxm, sxp);
                axis_bounds(by, ky, vs32, p.box_margin, p.smin32, sym, syp);
                axis_bounds(bz, kz, vs32, p.box_margin, p.smin32, szm, szp);
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
                    if (budget <= 0) {
                        heavy = true;
                        break;
                    }
                    visited |= 1u << nn;
                    const int ox = nn / 9 - 1, oy = (nn / 3) % 3 - 1, oz = nn % 3 - 1;
                    // an empty voxel costs one shared load here, so only occupied ones count against the budget
                    if (scan_slot(hs + ox * dyz + oy * dz + oz, bx - (float)ox * vs32, by - (float)oy * vs32, bz - (float)oz * vs32)) {
                        bound = prune_bound(p, min1);
                        --budget;
                    }
                }
                if (odd) heavy = false;  // met a record the f32 ranking cannot serve: exact path below
                if (!heavy && !odd && min1 < INF) {
                    if (min2 > band_limit(p, min1))
                        widx = idx1;
                    else
                        odd = true;
                }
            }
            TILE_STAMP(4);
            // ---- E: queries with many open neighbours, finished by their whole warp (lane = record) ---------------------------------
            {
                unsigned hm = __ballot_sync(FULL, heavy);
                while (hm) {
                    const int l = __ffs(hm) - 1;
                    hm &= hm - 1;
                    Carried cs;
                    cs.min1 = __shfl_sync(FULL, min1, l), cs.min2 = __shfl_sync(FULL, min2, l);
                    cs.idx1 = __shfl_sync(FULL, idx1, l), cs.visited = __shfl_sync(FULL, visited, l);
                    bool amb = false;
                    const uint32_t w = tile_finish_query_warp(p, sh, stage, lane, __shfl_sync(FULL, hs, l), dyz, dz, __shfl_sync(FULL, bx, l),
                                                              __shfl_sync(FULL, by, l), __shfl_sync(FULL, bz, l), __shfl_sync(FULL, kx, l),
                                                              __shfl_sync(FULL, ky, l), __shfl_sync(FULL, kz, l), __shfl_sync(FULL, qlf, l), cs, amb);
                    if (lane == l) widx = w, odd = amb;
                    if (COUNT && lane == 0) n_heavy += 1;
                }
            }
            TILE_STAMP(5);
            // ---- F: a unit whose region does not fit (never on a sorted scan): every query through the global-memory warp search ---
            if (!tiled) {
                unsigned fm = __ballot_sync(FULL, fast);
                unsigned long long d0 = 0, d1 = 0, d2 = 0;
                while (fm) {
                    const int l = __ffs(fm) - 1;
                    fm &= fm - 1;
                    const double4 c = ld256(p.src + c0 + warp * 32 + l);
                    const uint32_t w = search_query_warp<false>(p, lane, c, Carried{INF, INF, kNil, 0u}, d0, d1, d2);
                    if (lane == l) widx = w;
                }
            }
            // ---- G: queries the f32 ranking could not decide: the reference's own f64 sequence over all 27 voxels -------------------
            {
                unsigned em = __ballot_sync(FULL, valid && odd);
                while (em) {
                    const int l = __ffs(em) - 1;
                    em &= em - 1;
                    const double4 c = ld256(p.src + c0 + warp * 32 + l);
                    const uint32_t w = nn_exact<32>(p, FULL, lane, c.x, c.y, c.z, c.w, trunc_div(c.x, vs), trunc_div(c.y, vs), trunc_div(c.z, vs));
                    if (lane == l) widx = w;
                    if (COUNT && lane == 0) n_exact += 1;
                }
            }
            TILE_STAMP(6);
            // ---- H: acceptance on the exact records, residual, weight, the 16 sums + pair count --------------------------------------
            if (p.tgt_out != nullptr || __any_sync(FULL, widx != kNil)) {
                double a[kSums];
#pragma unroll
                for (int k = 0; k < kSums; ++k) a[k] = 0.0;
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
#pragma unroll
                for (int k = 0; k < kSums; ++k) {
                    a[k] += __shfl_xor_sync(FULL, a[k], 1);
                    a[k] += __shfl_xor_sync(FULL, a[k], 2);
                }
                if ((lane & 3) == 0) {
#pragma unroll
                    for (int k = 0; k < kSums; ++k) sh.acc[k][threadIdx.x >> 2] += a[k];
                }
            }
            TILE_STAMP(7);
            __syncthreads();  // (4) region table, staging area and boxes are rewritten by the next unit
            TILE_STAMP(8);
            if (p.dbg && threadIdx.x == 0) sh.dbg_t[10] += 1, sh.dbg_t[11] += uend - c0 < (uint32_t)kTileThreads ? uend - c0 : kTileThreads;
        }
    }
    if (p.dbg && threadIdx.x == 0) sh.dbg_t[9] = gtime();
#undef TILE_STAMP
    if (COUNT) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n_ranked += __shfl_xor_sync(FULL, n_ranked, o), n_probes += __shfl_xor_sync(FULL, n_probes, o);
            n_exact += __shfl_xor_sync(FULL, n_exact, o), n_heavy += __shfl_xor_sync(FULL, n_heavy, o);
            n_staged += __shfl_xor_sync(FULL, n_staged, o);
        }
        if (lane == 0) {
            atomicAdd(&st->stat_scanned, n_ranked), atomicAdd(&st->stat_probes, n_probes), atomicAdd(&st->stat_exact, n_exact);
            atomicAdd(&st->stat_heavy, n_heavy), atomicAdd(&st->stat_staged, n_staged);
        }
    }
    finish_iteration<kTileCols>(p, sh.acc, sh.est, sh.norm, sh.last, tag);
    if (p.dbg && threadIdx.x == 0)  // after finish_iteration, which stamps slot 3 for the per-query kernels
        for (int k = 0; k < kDbg; ++k) p.dbg[(size_t)kDbg * blockIdx.x + k] = sh.dbg_t[k];
}

// One launch = one Gauss-Newton iteration (`iteration0`: the first one of a registration).
// MINB: resident blocks per SM the register allocation aims at (6: 80 registers with some spills; 4: 128 registers, none)
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
    for (int i = 0; i < max_iterations; ++i) {
        nn_tile_iteration<false>(p, sh, s_tile_stage, i > 0, phase, p.xchg_tag + (unsigned long long)i);
        grid.sync();
        if (*reinterpret_cast<volatile int *>(&p.st->done)) break;
    }
}

}  // namespace sage
