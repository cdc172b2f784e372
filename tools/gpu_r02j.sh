#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tile.py -x -q --timeout=300 2>&1 | tail -3
TILE_TIMELINE=1 timeout 900 python tools/tile_probe.py legacy,tile 2000,15000,60000,120000 > gpurun_out/r02j_tile_probe.jsonl 2> gpurun_out/r02j_tile_probe.err; echo "rc=$?"
cut -c1-200 gpurun_out/r02j_tile_probe.jsonl; tail -5 gpurun_out/r02j_tile_probe.err
timeout 900 python bench.py --steps 20 --no-hbm-regime --no-pipeline > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r02j_bench.json
