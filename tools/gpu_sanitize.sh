#!/bin/bash
# compute-sanitizer over a small but complete slice: map build, both search schedules (tile search: persistent loop, one launch per
# iteration, tiny staging area, aliasing cells -> global fallback; per-query kernel: thread + warp modes), front end incl. the
# dynamic-vehicle filter and the key-frame grid, update, eviction
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import numpy as np, sys, os
sys.path.insert(0, '.')
import sage_icp_b200 as sg
from sage_icp_b200 import synthetic as syn
cfg = sg.launch_config(local_map_range=40.0, dynamic_vehicle_filter=True)
p = sg.SagePipeline(cfg)
traj = syn.trajectory(4)
for i in range(4):
    scan = syn.make_scan(i, tuple(traj[i]), n_beams=16, n_az=300)
    pose, _, _ = p.register_frame(scan)
g, ov = p.key_frame_grid(scan, [[-51.2, 51.2], [-51.2, 51.2], [-4, 2.4]], 128, 128)
print("pipeline ok", pose[:3], p.map().num_voxels(), g.sum())
pts = syn.sample_street_map(60000, 1, -20, 20)
scan = syn.make_scan(9, (0.0, 0.0, 0.0), n_beams=32, n_az=400)   # 12800 queries
guess = syn.pose7_from_xyyaw((0.2, 0.1, 0.004))
for name, env in (("per-query", {"SAGE_TILE": "0"}), ("tile persistent", {"SAGE_TILE_MIN": "1", "SAGE_TILE_FILL": "0"}),
                  ("tile per launch, 64 regs, tiny staging", {"SAGE_TILE_MIN": "1", "SAGE_TILE_FILL": "0", "SAGE_TILE_PERSISTENT": "0", "SAGE_TILE_MINB": "8", "SAGE_TILE_STAGE": "96"})):
    for k in ("SAGE_TILE", "SAGE_TILE_MIN", "SAGE_TILE_FILL", "SAGE_TILE_PERSISTENT", "SAGE_TILE_MINB", "SAGE_TILE_STAGE"):
        os.environ.pop(k, None)
    os.environ.update(env)
    m = sg.SageMap(0.8, 100.0, 20, 20, [40, 44, 48, 49, 50, 70, 72])
    m.add_points(pts)
    print(name, "big", m.register_frame(scan, guess, 3.0, 0.33, 0.4, max_iters=3, est_th=0.0)[0][:3])
    print(name, "small", m.register_frame(scan[:500], guess, 3.0, 0.33, 0.4, max_iters=3, est_th=0.0)[0][:3])
    far = scan[:2000].copy(); far[:, 0] += 409.6   # cells that alias in the sort key -> units that do not fit -> global fallback
    tgt, ok = m.get_correspondences(np.concatenate([scan[:3000], far]), 1.5, 0.4)
    print(name, "corr", ok.sum(), m.search_work(scan, 3.0, 0.4, with_staged=True), m.nn_stats(scan[:1000]))
m.remove_far([0, 0, 0]); print("voxels", m.num_voxels(), m.num_points())
PY
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --error-exitcode 7 python /tmp/san.py > gpurun_out/r02_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|pipeline ok|big|small|corr|voxels" gpurun_out/r02_sanitizer_$tool.log | cut -c1-160 | head -14
done
