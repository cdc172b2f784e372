#!/bin/bash
# visit AK: compute-sanitizer (racecheck, then memcheck) over the paths this round's second half changed — tile search with every
# block stepping and warp 0 publishing on its own, late draw, captured preparation graph (second and third call of one size), wide
# small-scan kernels with every block stepping, the warm-up launches
mkdir -p gpurun_out
cat > /tmp/san2.py <<'PY'
import numpy as np, sys, os
sys.path.insert(0, '.')
import sage_icp_b200 as sg
from sage_icp_b200 import synthetic as syn
pts = syn.sample_street_map(60000, 1, -20, 20)
scan = syn.make_scan(9, (0.0, 0.0, 0.0), n_beams=32, n_az=400)   # 12800 queries
guess = syn.pose7_from_xyyaw((0.2, 0.1, 0.004))
os.environ.update({"SAGE_TILE_MIN": "1", "SAGE_TILE_FILL": "0"})
m = sg.SageMap(0.8, 100.0, 20, 20, [40, 44, 48, 49, 50, 70, 72])
m.add_points(pts)
for k in range(3):  # plain launches, captured graph, replayed graph
    print("tile", k, m.register_frame(scan, guess, 3.0, 0.33, 0.4, max_iters=3, est_th=0.0)[0][:3])
os.environ.pop("SAGE_TILE_MIN")
m2 = sg.SageMap(0.8, 100.0, 20, 20, [40, 44, 48, 49, 50, 70, 72])
m2.add_points(pts)
for n in (500, 2000, 3000):  # 255-, 128- and 64-register instantiations of the small-scan loop
    print("small", n, m2.register_frame(scan[:n], guess, 3.0, 0.33, 0.4, max_iters=3, est_th=0.0)[0][:3])
PY
for tool in racecheck memcheck; do
  timeout 130 compute-sanitizer --tool $tool --error-exitcode 7 python /tmp/san2.py > gpurun_out/r02ak_sanitizer_$tool.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|tile|small" gpurun_out/r02ak_sanitizer_$tool.log | cut -c1-160 | head -10
done
