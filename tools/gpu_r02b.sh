#!/bin/bash
# round 2, visit B: where does the tile kernel's time go (phase timeline + ncu), dynamic-filter cluster order parity
mkdir -p gpurun_out
TILE_TIMELINE=1 timeout 600 python tools/tile_probe.py tile 120000 > gpurun_out/r02b_timeline.txt 2> gpurun_out/r02b_timeline.err; echo "rc=$?"
cat gpurun_out/r02b_timeline.txt | cut -c1-250; tail -3 gpurun_out/r02b_timeline.err
TILE_TIMELINE=1 timeout 600 python tools/tile_probe.py tile_probes27 120000 5000000 SAGE_TILE_PROBES=27 > gpurun_out/r02b_timeline27.txt 2>&1
cat gpurun_out/r02b_timeline27.txt | cut -c1-250
SAGE_TILE_PERSISTENT=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:nn_tile -s 12 -c 2 -f -o gpurun_out/r02b_tile \
    python tools/ncu_target.py > gpurun_out/r02b_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02b_ncu.log
timeout 1200 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_vs_reference_build.py -q --timeout=600 -k "dynamic" > gpurun_out/r02b_dynamic.log 2>&1; echo "rc=$?" >> gpurun_out/r02b_dynamic.log
tail -15 gpurun_out/r02b_dynamic.log
