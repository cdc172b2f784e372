#!/bin/bash
# round 2, visit V: log-free stop test, warp-0 publish, captured prep graph, wide small-scan kernels — parity subset + A/B timings
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tile.py tests/test_gpu_core.py tests/test_gpu_search_exactness.py tests/test_gpu_pipeline.py -x -q -m gpu \
    --timeout 150 --timeout-method=thread > gpurun_out/r02v_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02v_pytest.log; tail -4 gpurun_out/r02v_pytest.log
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-pipeline --no-hbm-regime > gpurun_out/r02v_bench_graph.json 2> gpurun_out/r02v_bench_graph.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/r02v_bench_graph.json
SAGE_TILE_GRAPH=0 timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-pipeline --no-hbm-regime > gpurun_out/r02v_bench_nograph.json 2> gpurun_out/r02v_bench_nograph.err; echo "bench rc=$?"
cut -c1-1500 gpurun_out/r02v_bench_nograph.json
timeout 300 python tools/small_probe.py > gpurun_out/r02v_small_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r02v_small_probe.log
