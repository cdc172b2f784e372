#!/bin/bash
# compute-sanitizer over a small but complete slice: map build, search kernel (thread + warp modes), front end, update, eviction
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import numpy as np, sys
sys.path.insert(0, '.')
import sage_icp_b200 as sg
from sage_icp_b200 import synthetic as syn
cfg = sg.launch_config(local_map_range=40.0)
p = sg.SagePipeline(cfg)
traj = syn.trajectory(4)
for i in range(4):
    pose, _, _ = p.register_frame(syn.make_scan(i, tuple(traj[i]), n_beams=16, n_az=300))
print("pipeline ok", pose[:3], p.map().num_voxels())
m = sg.SageMap(0.8, 100.0, 20, 20, [40, 44, 48, 49, 50, 70, 72])
m.add_points(syn.sample_street_map(60000, 1, -20, 20))
scan = syn.make_scan(9, (0.0, 0.0, 0.0), n_beams=32, n_az=400)   # 12800 queries: thread-per-query mode
guess = syn.pose7_from_xyyaw((0.2, 0.1, 0.004))
print("core big", m.register_frame(scan, guess, 3.0, 0.33, 0.4, max_iters=3, est_th=0.0))
print("core small", m.register_frame(scan[:500], guess, 3.0, 0.33, 0.4, max_iters=3, est_th=0.0))
tgt, ok = m.get_correspondences(scan[:3000], 1.5, 0.4)
print("corr", ok.sum(), m.search_work(scan, 3.0, 0.4), m.nn_stats(scan[:1000]))
m.remove_far([0, 0, 0]); print("voxels", m.num_voxels(), m.num_points())
PY
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python /tmp/san.py > gpurun_out/sanitizer_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|pipeline ok|core big|core small|corr|voxels" gpurun_out/sanitizer_$tool.log | head -10
done
