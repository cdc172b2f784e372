"""Scratch: kernel-level workload timing + search-work counters (not the contract bench)."""
import sys, os, time, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sage_icp_b200 as sg
import bench

n_map = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
half = bench.street_half_length(n_map)
pts = bench.make_map_points(n_map)
m = sg.SageMap(0.8, 1e9, 20, 20, bench.BASIC_LABELS)
m.add_points(pts)
print("map", m.num_points(), m.num_voxels())
for s in range(2):
    scan, guess = bench.make_queries(s, 64, 1875, half)
    yaw = 2.0 * math.atan2(guess[5], guess[6]); c, sn = math.cos(yaw), math.sin(yaw)
    qq = scan.copy()
    qq[:, 0] = c * scan[:, 0] - sn * scan[:, 1] + guess[0]; qq[:, 1] = sn * scan[:, 0] + c * scan[:, 1] + guess[1]; qq[:, 2] = scan[:, 2] + guess[2]
    occ, cand = m.nn_stats(qq)
    scanned, probes, exact, heavy = m.search_work(qq, 3.0, 0.4)
    n = len(qq)
    print(f"scan {s}: per query: occupied {occ/n:.2f} candidates {cand/n:.1f} | scanned {scanned/n:.1f} probes {probes/n:.2f} exact {exact} ({exact/n:.2e}) heavy {heavy} ({heavy/n:.3f})")
    for rep in range(3):
        m.profile_enable(True)
        t = time.time(); pose, it = m.register_frame(scan, guess, 3.0, 1/3, 0.4, max_iters=10, est_th=0.0); dt = time.time() - t
        nl, ms = m.profile_read()
        print(f"  rep {rep}: wall {dt*1e3:.2f} ms, nn kernel {nl} launches -> {ms/nl*1e3:.1f} us/iter")

# per-block timeline of one iteration
import ctypes as C
L = sg.load_library(); L.sage_debug_timeline.restype = C.c_size_t
n = L.sage_debug_timeline(m.h, None, C.c_size_t(0))
pose, it = m.register_frame(scan, guess, 3.0, 1/3, 0.4, max_iters=3, est_th=0.0)
buf = np.zeros(n, np.uint64)
L.sage_debug_timeline(m.h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_size_t(n))
K = 12
g = (n - 8) // K; b = buf[:K * g].reshape(g, K).astype(np.int64); tail = buf[K * g:K * g + 8].astype(np.int64)
t0 = b[:, 0].min()
pc = lambda a, q: np.percentile(a - t0, q)
print(f"timeline ns, grid {g}: start max {b[:,0].max()-t0} | warp0 light done med {pc(b[:,1],50):.0f} | block light done med {pc(b[:,2],50):.0f} p90 {pc(b[:,2],90):.0f} "
      f"max {pc(b[:,2],100):.0f} | block all done med {pc(b[:,3],50):.0f} p90 {pc(b[:,3],90):.0f} max {pc(b[:,3],100):.0f} | "
      f"last block: enter {tail[0]-t0} reduced {tail[1]-t0} | 6x6 solved {tail[5]-t0} exp done {tail[6]-t0} step done {tail[2]-t0}")
names = ["start", "loaded+transformed", "home probed", "home scanned", "neighbours done", "exact done", "accumulated"]
cols = [0, 4, 5, 6, 7, 8, 9]
prev = b[:, 0]
for nm, c in zip(names[1:], cols[1:]):
    d = b[:, c] - prev
    print(f"   warp0 phase {nm:20s}: median {np.median(d):8.0f} p90 {np.percentile(d,90):8.0f} max {d.max():8.0f} ns")
    prev = b[:, c]
