#!/bin/bash
# round 2, visit Y: the DRAM-bound regime (uniform queries, 20 M-point map) under the per-query kernel's runtime knobs
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 200 python tools/hbm_target.py 20000000 4 > gpurun_out/r02y_hbm_$name.json 2> gpurun_out/r02y_hbm_$name.err
  python -c "
import json
d=json.load(open('gpurun_out/r02y_hbm_$name.json')); print('$name', 'us/iter', round(d['us_per_iteration'],1), 'frac', round(d['frac'],3), 'launches', d['launches_timed'])" || tail -3 gpurun_out/r02y_hbm_$name.err; }
run default SAGE_X=0
run allwarp SAGE_ALL_WARP_MAX=100000000
run probes1 SAGE_LIGHT_PROBES=1
run probes3 SAGE_LIGHT_PROBES=3
run probes27 SAGE_LIGHT_PROBES=27
