"""Scratch: the per-iteration floor of pipeline-level clouds (a few hundred to a few thousand queries, nn_search_persistent_kernel)
on a 1 M-point map: us per Gauss-Newton iteration for the register-width variants (SAGE_SMALL_WIDE), then the per-block timeline of
one iteration (one launch per iteration: the debug stamps)."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sage_icp_b200 as sg
import bench

VARIANTS = {"default (255/128/64 regs, every block steps)": {}, "elected last block": {"SAGE_STEP_EVERYWHERE": "0"},
            "wide1 (128/64)": {"SAGE_SMALL_WIDE": "1"}, "wide0 (64)": {"SAGE_SMALL_WIDE": "0"}}
if os.environ.get("SMALL_PROBE_VARIANTS"):
    VARIANTS = {k: v for k, v in VARIANTS.items() if any(k.startswith(w) for w in os.environ["SMALL_PROBE_VARIANTS"].split(","))}
sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [300, 700, 1100, 2100, 5000, 12000]
n_map = 1_000_000
half = bench.street_half_length(n_map)
pts = bench.make_map_points(n_map)
scan, guess = bench.make_queries(0, 64, 1875, half)
for name, env in VARIANTS.items():
    os.environ.pop("SAGE_SMALL_WIDE", None)
    os.environ.pop("SAGE_STEP_EVERYWHERE", None)
    os.environ.update(env)
    m = sg.SageMap(0.8, 1e9, 20, 20, bench.BASIC_LABELS)
    m.add_points(pts)
    for n in sizes:
        sub = np.ascontiguousarray(scan[:: len(scan) // n][:n])
        best, wall_best, it = 1e9, 1e9, 0
        for rep in range(6):
            m.profile_enable(True)
            t = time.perf_counter()
            pose, it = m.register_frame(sub, guess, 3.0, 1 / 3, 0.4, max_iters=20, est_th=0.0)
            wall = time.perf_counter() - t
            nl, ms = m.profile_read()
            if rep:
                best, wall_best = min(best, ms / max(1, nl) * 1e3), min(wall_best, wall * 1e3)
        print(json.dumps({"variant": name, "n": len(sub), "us_per_iter": round(best, 2), "wall_ms_20_iters": round(wall_best, 3), "iters": it,
                          "pose": [float(f"{v:.12g}") for v in pose]}), flush=True)
    del m

import ctypes as C
os.environ.pop("SAGE_SMALL_WIDE", None)
os.environ.pop("SAGE_STEP_EVERYWHERE", None)
m = sg.SageMap(0.8, 1e9, 20, 20, bench.BASIC_LABELS)
m.add_points(pts)
L = sg.load_library(); L.sage_debug_timeline.restype = C.c_size_t
for n in (700, 2100):
    sub = np.ascontiguousarray(scan[:: len(scan) // n][:n])
    m.register_frame(sub, guess, 3.0, 1 / 3, 0.4, max_iters=2, est_th=0.0)
    nb = L.sage_debug_timeline(m.h, None, C.c_size_t(0))
    m.profile_enable(True)
    m.register_frame(sub, guess, 3.0, 1 / 3, 0.4, max_iters=4, est_th=0.0)
    nl, ms = m.profile_read()
    buf = np.zeros(nb, np.uint64)
    L.sage_debug_timeline(m.h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_size_t(nb))
    K = 12
    grid = (n + 7) // 8
    b = buf[:K * grid].reshape(grid, K).astype(np.int64)
    tail = buf[K * grid:K * grid + 8].astype(np.int64)
    t0 = b[:, 0].min()
    print(f"n={n}: one launch per iteration {ms / max(1, nl) * 1e3:.1f} us/launch; grid {grid} (tail says {tail[3]}); block start spread {b[:,0].max()-t0} ns; "
          f"search done (finish entry) med {np.median(b[:,3]-t0):.0f} p90 {np.percentile(b[:,3]-t0,90):.0f} max {(b[:,3]-t0).max()} ns; "
          f"last block: enters reduce {tail[0]-t0}, reduced +{tail[1]-tail[0]}, solved +{tail[5]-tail[1]}, exp +{tail[6]-tail[5]}, step done +{tail[2]-tail[6]} "
          f"(end of kernel work {tail[2]-t0} ns)")
