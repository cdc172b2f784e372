#!/bin/bash
# scratch: staged-scan experiment
mkdir -p gpurun_out
timeout 200 python tools/stage_probe.py 0,16,40,8 2>&1 | tail -12 | tee gpurun_out/stage_probe.log
SAGE_STAGE_CAP=16 timeout 200 python -m pytest tests/test_gpu_search_exactness.py tests/test_gpu_core.py -x -q -m gpu --timeout 120 --timeout-method=thread 2>&1 | tail -5 | tee gpurun_out/stage_parity16.log
