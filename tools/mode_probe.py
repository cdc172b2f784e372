"""Scratch: search-kernel time vs number of queries, thread-per-query mode vs warp-per-query mode."""
import sys, os, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sage_icp_b200 as sg
import bench
half = bench.street_half_length(5_000_000)
m = sg.SageMap(0.8, 1e9, 20, 20, bench.BASIC_LABELS)
m.add_points(bench.make_map_points(5_000_000))
scan, guess = bench.make_queries(0, 64, 1875, half)
for n in (2000, 5000, 10000, 15000, 30000, 60000, 120000):
    sub = np.ascontiguousarray(scan[:: len(scan) // n][:n])
    for rep in range(3):
        m.profile_enable(True)
        pose, it = m.register_frame(sub, guess, 3.0, 1/3, 0.4, max_iters=10, est_th=0.0)
        nl, ms = m.profile_read()
    print(f"n={len(sub):7d}: {ms/nl*1e3:7.1f} us/iter")
