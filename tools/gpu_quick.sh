#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_core.py tests/test_gpu_search_exactness.py tests/test_golden.py tests/test_gpu_fuzz.py -m gpu -x -q --timeout=60 --timeout-method=thread 2>&1 | tail -3
timeout 200 python tools/perf_probe.py 2>&1 | grep -E "scan 0|rep 2|timeline|phase"
