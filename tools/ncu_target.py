"""Scratch: a few kernel-level registrations on BASELINE configs[1] for ncu to attach to (the search schedule comes from SAGE_*)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sage_icp_b200 as sg
import bench
n_map = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
half = bench.street_half_length(n_map)
m = sg.SageMap(0.8, 1e9, 20, 20, bench.BASIC_LABELS)
m.add_points(bench.make_map_points(n_map))
scan, guess = bench.make_queries(0, 64, 1875, half)
for r in range(reps):
    pose, it = m.register_frame(scan, guess, 3.0, 1 / 3, 0.4, max_iters=10, est_th=0.0)
print(pose, it)
