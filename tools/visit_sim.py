"""CPU model of the search kernel's voxel-visit schedule on the bench workload (BASELINE configs[1]) — no GPU needed.

For every query it rebuilds, from the oracle's map, the lower bound of the metric over each of the 27 neighbour boxes and the best
metric actually stored in each voxel, replays the kernel's nearest-box-first search with the prune bound re-tightened after every
visit, and reports (a) visits per query (cross-check: the device counter `table_probes` in bench.py's kernel_work_per_query),
(b) what a warp of 32 consecutive queries pays today = the visits of its slowest lane, and (c) what it would pay if the pending
visits of the 32 lanes were pooled and dealt back evenly each round.  Test infrastructure: uses
oracle/ for the map, like bench.py's cpu_baseline leg.

    python tools/visit_sim.py [--map-points 5000000] [--queries 120000]
"""
import argparse, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench


def box_gap(x, k_home, off, vs):
    """distance from coordinate x (home key k_home) to the box of key k_home + off along one axis (truncation-toward-zero keys)"""
    k = k_home + off
    lo = np.where(k > 0, k * vs, np.where(k == 0, -vs, (k - 1) * vs))
    hi = np.where(k > 0, (k + 1) * vs, np.where(k == 0, vs, k * vs))
    return np.maximum(np.maximum(lo - x, x - hi), 0.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--map-points", type=int, default=5_000_000)
    ap.add_argument("--queries", type=int, default=120_000)
    ap.add_argument("--chunk", type=int, default=4096)
    a = ap.parse_args()
    from oracle import oracle_py as orc
    vs, th = bench.VOXEL_SIZE_MAP, bench.SEM_TH
    smin = min(1.0, th)
    t0 = time.time()
    om = orc.OracleMap(vs, 1e9, bench.BASIC, bench.CRITICAL, bench.BASIC_LABELS, evict_faithful=False)
    om.add_points(bench.make_map_points(a.map_points))
    keys, counts, pts = om.dump()
    half = bench.street_half_length(a.map_points)
    scan, guess = bench.make_queries(0, 64, 1875, half)
    yaw = 2.0 * math.atan2(guess[5], guess[6]); c, s = math.cos(yaw), math.sin(yaw)
    q = scan.copy()
    q[:, 0] = c * scan[:, 0] - s * scan[:, 1] + guess[0]; q[:, 1] = s * scan[:, 0] + c * scan[:, 1] + guess[1]; q[:, 2] = scan[:, 2] + guess[2]
    q = q[: a.queries]
    print(f"map {counts.sum()} points / {len(keys)} voxels, {len(q)} queries ({time.time() - t0:.1f} s)", flush=True)

    B = 1 << 21
    pack = lambda k: (k[..., 0].astype(np.int64) + B) | ((k[..., 1].astype(np.int64) + B) << 22) | ((k[..., 2].astype(np.int64) + B) << 44)
    pk = pack(keys)
    order = np.argsort(pk); pk = pk[order]; counts = counts[order]; pts = pts[order]
    offs = np.array([[ox, oy, oz] for ox in (-1, 0, 1) for oy in (-1, 0, 1) for oz in (-1, 0, 1)])  # the reference's enumeration
    INF = np.inf
    n = len(q)
    visits = np.zeros(n, np.int32)      # voxel probes per query, home included
    useful = np.zeros(n, np.int32)      # of which found a non-empty voxel
    pooled_rounds, warp_max, warp_sum = [], [], []
    for c0 in range(0, n, a.chunk):
        qq = q[c0:c0 + a.chunk]; m = len(qq)
        kh = np.trunc(qq[:, :3] / vs).astype(np.int64)
        # lower bound of the metric over each neighbour box
        gap2 = np.zeros((m, 27))
        for ax in range(3):
            gap2 += box_gap(qq[:, ax:ax + 1], kh[:, ax:ax + 1], offs[None, :, ax], vs) ** 2
        lb = smin * gap2
        # best stored metric per neighbour voxel
        nk = kh[:, None, :] + offs[None, :, :]
        pos = np.searchsorted(pk, pack(nk)); pos = np.minimum(pos, len(pk) - 1)
        found = pk[pos] == pack(nk)
        best = np.full((m, 27), INF)
        qi, vi = np.nonzero(found)
        P = pts[pos[qi, vi]]                                   # (items, stride, 4)
        cnt = counts[pos[qi, vi]]
        d = ((P[:, :, :3] - qq[qi, None, :3]) ** 2).sum(-1)
        ln, lq = np.trunc(P[:, :, 3]), np.trunc(qq[qi, None, 3])
        compat = (ln == lq) | (np.trunc(P[:, :, 3] * qq[qi, None, 3]) == 0)
        metric = np.where(compat, d * th, d)
        metric[np.arange(P.shape[1])[None, :] >= cnt[:, None]] = INF
        best[qi, vi] = metric.min(1)
        nonempty = np.isfinite(best)
        # today's schedule: home, then nearest open box first, bound re-tightened after every visit
        visited = np.zeros((m, 27), bool); visited[:, 13] = True
        min1 = best[:, 13].copy()
        nv = np.ones(m, np.int32); nu = nonempty[:, 13].astype(np.int32)
        seq = [np.full(m, 13)]
        while True:
            cand = np.where(~visited & (lb <= min1[:, None]), lb, INF)
            j = cand.argmin(1); go = np.isfinite(cand[np.arange(m), j])
            if not go.any():
                break
            r = np.nonzero(go)[0]
            visited[r, j[r]] = True
            min1[r] = np.minimum(min1[r], best[r, j[r]])
            nv[r] += 1; nu[r] += nonempty[r, j[r]]
        visits[c0:c0 + m] = nv; useful[c0:c0 + m] = nu
        assert np.array_equal(min1, best.min(1))  # pruned search == scan of all 27 voxels
        # per warp of 32 consecutive queries
        for w0 in range(0, m - m % 32, 32):
            sl = slice(w0, w0 + 32)
            warp_max.append(nv[sl].max()); warp_sum.append(nv[sl].sum())
            # pooled schedule: after the home voxel, every round deals up to 32 (query, voxel) visits: each active query nominates
            # its nearest open boxes, floor(32 / active) of them (at least one); bounds re-tightened between rounds only
            v2 = np.zeros((32, 27), bool); v2[:, 13] = True
            m1 = best[sl, 13].copy(); l2 = lb[sl]; b2 = best[sl]
            rounds = 0
            while True:
                open_ = ~v2 & (l2 <= m1[:, None])
                act = open_.any(1)
                if not act.any():
                    break
                quota = max(1, 32 // int(act.sum()))
                rounds += 1
                for r in np.nonzero(act)[0]:
                    idx = np.nonzero(open_[r])[0]
                    idx = idx[np.argsort(l2[r, idx], kind="stable")][:quota]
                    v2[r, idx] = True
                    m1[r] = min(m1[r], b2[r, idx].min())
            pooled_rounds.append(rounds)
            assert np.array_equal(m1, b2.min(1))  # so does the pooled schedule
        print(f"  {min(c0 + a.chunk, n)} queries", end="\r", flush=True)
    warp_max, warp_sum, pooled_rounds = map(np.array, (warp_max, warp_sum, pooled_rounds))
    print()
    print(f"visits per query (home included): mean {visits.mean():.2f}, p50 {np.percentile(visits, 50):.0f}, p90 {np.percentile(visits, 90):.0f}, "
          f"p99 {np.percentile(visits, 99):.0f}, max {visits.max()};  visits that find a non-empty voxel: {useful.mean():.2f}")
    print(f"per warp of 32 consecutive queries: slowest lane (what the thread-per-query phase pays, before any deferral) mean {warp_max.mean():.2f}, "
          f"p50 {np.percentile(warp_max, 50):.0f}, p90 {np.percentile(warp_max, 90):.0f}, max {warp_max.max()};  sum of visits / 32 = {warp_sum.mean() / 32:.2f}")
    print(f"pooled schedule: 1 home round + neighbour rounds: mean {1 + pooled_rounds.mean():.2f}, p50 {1 + np.percentile(pooled_rounds, 50):.0f}, "
          f"p90 {1 + np.percentile(pooled_rounds, 90):.0f}, max {1 + pooled_rounds.max()}")


if __name__ == "__main__":
    main()
