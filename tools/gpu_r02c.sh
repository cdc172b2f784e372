#!/bin/bash
# round 2, visit C: tile search v2 (rounds 0/1 by the owner thread, round 2 pooled, dynamic units) — parity, timing, timeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tile.py -x -q --timeout=200 > gpurun_out/r02c_tile_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02c_tile_tests.log
tail -8 gpurun_out/r02c_tile_tests.log
TILE_TIMELINE=1 timeout 900 python tools/tile_probe.py > gpurun_out/r02c_tile_probe.jsonl 2> gpurun_out/r02c_tile_probe.err; echo "rc=$?"
cut -c1-220 gpurun_out/r02c_tile_probe.jsonl; tail -5 gpurun_out/r02c_tile_probe.err
timeout 1200 python -m pytest tests/test_gpu_search_exactness.py tests/test_gpu_fuzz.py tests/test_gpu_core.py -q -x --timeout=300 > gpurun_out/r02c_exact.log 2>&1; echo "rc=$?" >> gpurun_out/r02c_exact.log
tail -8 gpurun_out/r02c_exact.log
SAGE_TILE_PERSISTENT=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:nn_tile -s 12 -c 1 -f -o gpurun_out/r02c_tile \
    python tools/ncu_target.py > gpurun_out/r02c_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r02c_ncu.log
