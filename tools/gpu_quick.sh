#!/bin/bash
# One short GPU-box visit: the fast parity subset (core, exactness, pipeline), then the smoke.  For the full round use tools/gpu_round.sh.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_core.py tests/test_gpu_search_exactness.py tests/test_gpu_pipeline.py tests/test_gpu_vs_reference_build.py -x -q -m gpu \
    --timeout 150 --timeout-method=thread 2>&1 | tail -8 | tee gpurun_out/quick.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.log
