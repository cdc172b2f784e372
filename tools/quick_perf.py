"""Scratch timing of the kernel-level workload (not the contract bench): map build + 10 GN iterations."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sage_icp_b200 as sg
from sage_icp_b200 import synthetic as syn

n_map = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
L = 0.36e-3 * n_map / 2  # street half-length so that voxels stay ~full
t = time.time(); pts = syn.sample_street_map(n_map * 2, 1, -L, L); print("map gen", time.time() - t, len(pts))
m = sg.SageMap(0.8, 1e9, 20, 20, [40, 44, 48, 49, 50, 70, 72])
t = time.time(); m.add_points(pts); print("gpu add_points", time.time() - t, "voxels", m.num_voxels(), "points", m.num_points())
scan = syn.make_scan(0, (0.0, 0.0, 0.0))
q = scan.copy(); q[:, 2] += syn.SENSOR_HEIGHT
r = np.linalg.norm(scan[:, :3], axis=1); q = q[(r > 5) & (r < 100)]
print("queries", len(q), "stats", m.nn_stats(q), "per query", np.array(m.nn_stats(q)) / len(q))
guess = np.array([0.3, 0.1, 0.0, 0, 0, np.sin(0.005), np.cos(0.005)])
for rep in range(3):
    m.profile_enable(True)
    t = time.time(); pose, it = m.register_frame(q, guess, 3.0, 0.33, 0.4, max_iters=10, est_th=0.0); dt = time.time() - t
    nl, ms = m.profile_read()
    print(f"rep {rep}: {it} iters, wall {dt*1e3:.2f} ms, nn kernel {nl} launches {ms:.3f} ms total -> {ms/nl*1e3:.1f} us/iter", pose[:3])
