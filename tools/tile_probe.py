"""Scratch: A/B timing of the search schedules on BASELINE configs[1] (5 M-point map, 120 k-query scan and sub-samples).
Each configuration is a set of SAGE_* variables the library reads when a map first searches, so each gets its own map."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sage_icp_b200 as sg
import bench

CONFIGS = {
    "legacy": {"SAGE_TILE": "0"},
    "tile": {"SAGE_TILE_MIN": "1"},
    "tile_launch_per_iter": {"SAGE_TILE_MIN": "1", "SAGE_TILE_PERSISTENT": "0"},
    "tile_minb4": {"SAGE_TILE_MIN": "1", "SAGE_TILE_MINB": "4"},
    "tile_minb6_stage1024": {"SAGE_TILE_MIN": "1", "SAGE_TILE_MINB": "6", "SAGE_TILE_STAGE": "1024"},
    "tile_minb8_stage704": {"SAGE_TILE_MIN": "1", "SAGE_TILE_MINB": "8", "SAGE_TILE_STAGE": "704"},
    "tile_minb6_stage1408": {"SAGE_TILE_MIN": "1", "SAGE_TILE_MINB": "6", "SAGE_TILE_STAGE": "1408"},
    "tile_minb4_stage2560": {"SAGE_TILE_MIN": "1", "SAGE_TILE_MINB": "4", "SAGE_TILE_STAGE": "2560"},
    "tile_minb5_stage1024": {"SAGE_TILE_MIN": "1", "SAGE_TILE_MINB": "5", "SAGE_TILE_STAGE": "1024"},
    "tile_stage2048": {"SAGE_TILE_MIN": "1", "SAGE_TILE_STAGE": "2048"},
    "tile_list_order": {"SAGE_TILE_MIN": "1", "SAGE_TILE_BY_SIZE": "0"},
    "tile_blocks4": {"SAGE_TILE_MIN": "1", "SAGE_TILE_BLOCKS": "4"},
    "tile_blocks3": {"SAGE_TILE_MIN": "1", "SAGE_TILE_BLOCKS": "3"},
    "tile_elected": {"SAGE_TILE_MIN": "1", "SAGE_STEP_EVERYWHERE": "0"},
    "tile_every_block": {"SAGE_TILE_MIN": "1", "SAGE_STEP_EVERYWHERE": "2"},
}
KEYS = ("SAGE_TILE", "SAGE_TILE_BY_SIZE", "SAGE_TILE_FILL", "SAGE_TILE_MIN", "SAGE_TILE_PERSISTENT", "SAGE_TILE_MINB", "SAGE_TILE_STAGE", "SAGE_TILE_BLOCKS", "SAGE_STEP_EVERYWHERE")
which = sys.argv[1].split(",") if len(sys.argv) > 1 else list(CONFIGS)
sizes = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [2000, 8000, 15000, 30000, 60000, 120000]
n_map = int(sys.argv[3]) if len(sys.argv) > 3 else 5_000_000
half = bench.street_half_length(n_map)
pts = bench.make_map_points(n_map)
scan, guess = bench.make_queries(0, 64, 1875, half)
rows = []
for name in which:
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(CONFIGS[name])
    if CONFIGS[name].get("SAGE_TILE_MIN") == "1":
        os.environ["SAGE_TILE_FILL"] = "0"
    m = sg.SageMap(0.8, 1e9, 20, 20, bench.BASIC_LABELS)
    m.add_points(pts)
    for n in sizes:
        sub = np.ascontiguousarray(scan[:: len(scan) // n][:n])
        best_it, best_wall, pose = 1e9, 1e9, None
        for rep in range(5):
            m.profile_enable(True)
            t = time.perf_counter()
            pose, it = m.register_frame(sub, guess, 3.0, 1 / 3, 0.4, max_iters=10, est_th=0.0)
            wall = time.perf_counter() - t
            nl, ms = m.profile_read()
            if rep:
                best_it, best_wall = min(best_it, ms / max(1, nl) * 1e3), min(best_wall, wall * 1e3)
        row = {"config": name, "n": len(sub), "us_per_iter": round(best_it, 2), "wall_ms": round(best_wall, 3), "iters": it,
               "pose": [float(f"{v:.12g}") for v in pose]}
        if n == sizes[-1]:
            w = m.search_work(sub, 3.0, 0.4, with_staged=True)
            row["work_per_query"] = {"ranked": w[0] / len(sub), "probes": w[1] / len(sub), "exact": w[2] / len(sub), "pooled_pairs": w[3] / len(sub),
                                     "staged": w[4] / len(sub)}
        rows.append(row)
        print(json.dumps(row), flush=True)
    del m

# per-block phase timeline of the tile kernel (one launch per iteration; the last of 3 iterations is what the buffer holds)
if os.environ.get("TILE_TIMELINE", "1") != "0":
    import ctypes as C
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update({"SAGE_TILE_MIN": "1", "SAGE_TILE_FILL": "0"})
    if len(sys.argv) > 4:
        os.environ.update(dict(kv.split("=") for kv in sys.argv[4].split(",")))
    m = sg.SageMap(0.8, 1e9, 20, 20, bench.BASIC_LABELS)
    m.add_points(pts)
    L = sg.load_library(); L.sage_debug_timeline.restype = C.c_size_t
    if os.environ.get("TILE_TIMELINE_N"):
        tn = int(os.environ["TILE_TIMELINE_N"])
        scan = np.ascontiguousarray(scan[:: len(scan) // tn][:tn])
    m.register_frame(scan, guess, 3.0, 1 / 3, 0.4, max_iters=2, est_th=0.0)
    n = L.sage_debug_timeline(m.h, None, C.c_size_t(0))
    m.register_frame(scan, guess, 3.0, 1 / 3, 0.4, max_iters=3, est_th=0.0)
    buf = np.zeros(n, np.uint64)
    L.sage_debug_timeline(m.h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), C.c_size_t(n))
    K = 12
    g = (n - 8) // K
    b = buf[:K * g].reshape(g, K).astype(np.int64)
    b = b[b[:, 0] > 0]
    t0 = b[:, 0].min()
    names = {8: "unit fetch", 1: "A load+transform+box", 2: "C probes+scan+issue", 3: "bulk wait", 4: "D rounds 0+1 (own thread)", 5: "E round 2 (pooled)",
             6: "F/G fallback+exact", 7: "H accept+sums"}
    print(f"tile timeline: {len(b)} blocks, start spread {b[:,0].max()-t0} ns, block end (before finish) med {np.median(b[:,9]-t0):.0f} p90 "
          f"{np.percentile(b[:,9]-t0,90):.0f} max {(b[:,9]-t0).max()} ns; units/block med {np.median(b[:,10]):.1f} max {b[:,10].max()}, "
          f"queries/block med {np.median(b[:,11]):.0f} max {b[:,11].max()}")
    order_ = np.argsort(-(b[:, 9] - t0))[:8]
    print("   slowest blocks: end | units queries | fetch A C wait D E FG H (ns, summed over the block's units)")
    for i_ in order_:
        print(f"      {b[i_,9]-t0:7d} | {b[i_,10]:2d} {b[i_,11]:4d} | " + " ".join(f"{b[i_,k_]:6d}" for k_ in (8, 1, 2, 3, 4, 5, 6, 7)))
    tail = buf[K * g:K * g + 8].astype(np.int64)
    print(f"   last block: enters reduce {tail[0]-t0} ns, sums reduced/exchanged +{tail[1]-tail[0]}, 6x6 solved +{tail[5]-tail[1]}, exp +{tail[6]-tail[5]}, "
          f"step done +{tail[2]-tail[6]} (total {tail[2]-tail[0]} ns)")
    for k, nm in names.items():
        print(f"   {nm:24s}: per block total ns: median {np.median(b[:,k]):8.0f} p90 {np.percentile(b[:,k],90):8.0f} max {b[:,k].max():8d} | "
              f"per unit mean {b[:,k].sum()/max(1,b[:,10].sum()):7.0f}")
