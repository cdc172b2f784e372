#!/bin/bash
# round 2, visit D: tile search v3 (merged units, prefetch) + the new bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tile.py -x -q --timeout=200 > gpurun_out/r02d_tile_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02d_tile_tests.log
tail -5 gpurun_out/r02d_tile_tests.log
TILE_TIMELINE=1 timeout 900 python tools/tile_probe.py legacy,tile,tile_minb4,tile_minb6_stage1024,tile_minb5_stage1024 > gpurun_out/r02d_tile_probe.jsonl 2> gpurun_out/r02d_tile_probe.err; echo "rc=$?"
cut -c1-200 gpurun_out/r02d_tile_probe.jsonl; tail -5 gpurun_out/r02d_tile_probe.err
timeout 900 python bench.py --steps 20 > gpurun_out/r02d_bench.json 2> gpurun_out/r02d_bench.err; echo "bench rc=$?"; cat gpurun_out/r02d_bench.json; tail -5 gpurun_out/r02d_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02d_bench_reference.json 2> gpurun_out/r02d_bench_reference.err; cat gpurun_out/r02d_bench_reference.json
