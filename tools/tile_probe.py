"""Scratch: A/B timing of the search schedules on BASELINE configs[1] (5 M-point map, 120 k-query scan and sub-samples).
Each configuration is a set of SAGE_* variables the library reads when a map first searches, so each gets its own map."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import sage_icp_b200 as sg
import bench

CONFIGS = {
    "legacy": {"SAGE_TILE": "0"},
    "legacy_pooled": {"SAGE_TILE": "0", "SAGE_POOLED": "1"},
    "tile": {"SAGE_TILE_MIN": "1"},
    "tile_launch_per_iter": {"SAGE_TILE_MIN": "1", "SAGE_TILE_PERSISTENT": "0"},
    "tile_128regs": {"SAGE_TILE_MIN": "1", "SAGE_TILE_MINB": "4"},
    "tile_probes0": {"SAGE_TILE_MIN": "1", "SAGE_TILE_PROBES": "0"},
    "tile_probes1": {"SAGE_TILE_MIN": "1", "SAGE_TILE_PROBES": "1"},
    "tile_probes8": {"SAGE_TILE_MIN": "1", "SAGE_TILE_PROBES": "8"},
    "tile_probes27": {"SAGE_TILE_MIN": "1", "SAGE_TILE_PROBES": "27"},
    "tile_stage1024": {"SAGE_TILE_MIN": "1", "SAGE_TILE_STAGE": "1024"},
    "tile_stage2560": {"SAGE_TILE_MIN": "1", "SAGE_TILE_STAGE": "2560"},
    "tile_blocks4": {"SAGE_TILE_MIN": "1", "SAGE_TILE_BLOCKS": "4"},
    "tile_blocks3": {"SAGE_TILE_MIN": "1", "SAGE_TILE_BLOCKS": "3"},
}
KEYS = ("SAGE_TILE", "SAGE_POOLED", "SAGE_TILE_MIN", "SAGE_TILE_PERSISTENT", "SAGE_TILE_MINB", "SAGE_TILE_PROBES", "SAGE_TILE_STAGE", "SAGE_TILE_BLOCKS")
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
            row["work_per_query"] = {"ranked": w[0] / len(sub), "probes": w[1] / len(sub), "exact": w[2] / len(sub), "warp_phase": w[3] / len(sub),
                                     "staged": w[4] / len(sub)}
        rows.append(row)
        print(json.dumps(row), flush=True)
    del m
