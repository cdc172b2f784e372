#!/usr/bin/env python
"""BASELINE configs[4]: map-size sweep at a fixed 120 k-point query set — search-kernel time and algorithmic GB/s against
the HBM roofline, 1 M -> 50 M map points.  Two query sets per map: (a) one synthetic scan (spatially coherent: the touched
part of the map stays L2-resident whatever the map size), (b) 120 k queries spread uniformly over the whole map (no reuse:
this is what exposes the DRAM-bound regime).  Prints one JSON line per (map, query set).  Not the contract bench."""
import argparse
import json
import math
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="1,5,20,50")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import torch
    import sage_icp_b200 as sg
    from sage_icp_b200 import synthetic as syn
    peak, peak_src = bench.measured_peak_gbs()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for mp in [int(float(x) * 1e6) for x in a.sizes.split(",")]:
        half = bench.street_half_length(mp)
        m = sg.SageMap(bench.VOXEL_SIZE_MAP, 1e9, bench.BASIC, bench.CRITICAL, bench.BASIC_LABELS)
        want, made, seed, t0 = int(2.17 * mp), 0, 1000, time.time()
        rng = np.random.default_rng(7)
        uniform_q = []
        while made < want:  # generate + insert chunk by chunk: no multi-GB host array
            k = min(4_000_000, want - made)
            pts = syn.sample_street_map(k, seed, -half, half)
            m.add_points(pts)
            uniform_q.append(pts[rng.choice(len(pts), max(1, int(120_000 * k / want)), replace=False)])
            made += k; seed += 1
        build_s = time.time() - t0
        n_pts, n_vox = m.num_points(), m.num_voxels()
        scan, guess = bench.make_queries(0, 64, 1875, half)
        uq = np.concatenate(uniform_q)[:120_000].copy()
        uq[:, :3] += rng.normal(0, 0.1, (len(uq), 3))  # near, not on, map points
        ident = np.array([0, 0, 0, 0, 0, 0, 1.0])
        for name, q, g in (("scan", scan, guess), ("uniform", uq, ident)):
            yaw = 2.0 * math.atan2(g[5], g[6]); c, s = math.cos(yaw), math.sin(yaw)
            qq = q.copy()
            qq[:, 0] = c * q[:, 0] - s * q[:, 1] + g[0]; qq[:, 1] = s * q[:, 0] + c * q[:, 1] + g[1]; qq[:, 2] = q[:, 2] + g[2]
            occ, cand = m.nn_stats(qq)
            work = m.search_work(qq, bench.MAX_DIST, bench.SEM_TH)
            alg = bench.algorithmic_bytes(len(qq), occ, cand)
            d = torch.from_numpy(np.ascontiguousarray(q)).cuda()
            for _ in range(2):
                m.register_frame_device(d.data_ptr(), len(q), g, bench.MAX_DIST, bench.KERNEL, bench.SEM_TH, 10, 0.0)
            m.profile_enable(True)
            for _ in range(a.reps):
                flush.fill_(1); torch.cuda.synchronize()
                m.register_frame_device(d.data_ptr(), len(q), g, bench.MAX_DIST, bench.KERNEL, bench.SEM_TH, 10, 0.0)
            nl, ms = m.profile_read()
            m.profile_enable(False)
            us = 1e3 * ms / nl
            print(json.dumps({"map_points": n_pts, "map_voxels": n_vox, "map_build_s": round(build_s, 1), "queries": name, "n_queries": len(q),
                              "occupied_voxels_per_query": occ / len(q), "candidates_per_query": cand / len(q),
                              "records_scanned_per_query": work[0] / len(q), "probes_per_query": work[1] / len(q),
                              "algorithmic_bytes_per_launch": alg, "kernel_us": us, "achieved_gbs": alg / us / 1e3,
                              "frac_of_peak": alg / us / 1e3 / peak, "peak_gbs": peak, "peak_source": peak_src}), flush=True)
        del m
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
