#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tile.py -x -q --timeout=300 2>&1 | tail -2
TILE_TIMELINE=1 timeout 900 python tools/tile_probe.py legacy,tile,tile_list_order 15000,60000,120000 > gpurun_out/r02q_tile_probe.jsonl 2> gpurun_out/r02q_tile_probe.err; echo "rc=$?"
cut -c1-120 gpurun_out/r02q_tile_probe.jsonl | head -30; tail -3 gpurun_out/r02q_tile_probe.err
timeout 900 python bench.py --steps 20 --no-hbm-regime --no-pipeline --no-cpu-baseline > gpurun_out/r02q_bench.json 2> gpurun_out/r02q_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02q_bench.json
SAGE_TILE_BY_SIZE=0 timeout 900 python bench.py --steps 20 --no-hbm-regime --no-pipeline --no-cpu-baseline > gpurun_out/r02q_bench_listorder.json 2> gpurun_out/r02q_bench_listorder.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r02q_bench_listorder.json
