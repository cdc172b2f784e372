"""Scratch experiment: does feeding the queries in voxel order (host-sorted here) speed the search kernel up?"""
import sys, os, time, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sage_icp_b200 as sg
import bench

n_map = 5_000_000
half = bench.street_half_length(n_map)
m = sg.SageMap(0.8, 1e9, 20, 20, bench.BASIC_LABELS)
m.add_points(bench.make_map_points(n_map))
scan, guess = bench.make_queries(0, 64, 1875, half)
yaw = 2.0 * math.atan2(guess[5], guess[6]); c, s = math.cos(yaw), math.sin(yaw)
qq = scan.copy()
qq[:, 0] = c * scan[:, 0] - s * scan[:, 1] + guess[0]; qq[:, 1] = s * scan[:, 0] + c * scan[:, 1] + guess[1]; qq[:, 2] = scan[:, 2] + guess[2]
k = np.trunc(qq[:, :3] / 0.8).astype(np.int64)
def morton(k):
    k = k - k.min(0)
    out = np.zeros(len(k), np.int64)
    for b in range(11):
        for a in range(3):
            out |= ((k[:, a] >> b) & 1) << (3 * b + a)
    return out
orders = {"scan order": np.arange(len(scan)), "voxel key (lexsort x,y,z)": np.lexsort((k[:, 2], k[:, 1], k[:, 0])),
          "voxel key morton": np.argsort(morton(k), kind="stable"), "random": np.random.default_rng(0).permutation(len(scan))}
for name, o in orders.items():
    sc = np.ascontiguousarray(scan[o])
    for rep in range(3):
        m.profile_enable(True)
        pose, it = m.register_frame(sc, guess, 3.0, 1/3, 0.4, max_iters=10, est_th=0.0)
        nl, ms = m.profile_read()
    print(f"{name:28s}: {ms/nl*1e3:7.1f} us/iter  pose {pose[:3]}")
