#!/bin/bash
# round 2, visit A: first run of the tile search — parity first, then A/B timing of the schedules
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_tile.py -x -q --timeout=200 > gpurun_out/r02a_tile_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02a_tile_tests.log
tail -15 gpurun_out/r02a_tile_tests.log
timeout 1200 python -m pytest tests/test_gpu_search_exactness.py tests/test_gpu_fuzz.py tests/test_gpu_core.py -q --timeout=300 > gpurun_out/r02a_exact.log 2>&1; echo "rc=$?" >> gpurun_out/r02a_exact.log
tail -15 gpurun_out/r02a_exact.log
timeout 900 python tools/tile_probe.py > gpurun_out/r02a_tile_probe.jsonl 2> gpurun_out/r02a_tile_probe.err; echo "rc=$?"
cat gpurun_out/r02a_tile_probe.jsonl | cut -c1-200
tail -5 gpurun_out/r02a_tile_probe.err
