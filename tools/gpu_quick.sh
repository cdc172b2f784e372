#!/bin/bash
timeout 400 python -m pytest tests/test_gpu_core.py tests/test_gpu_search_exactness.py tests/test_golden.py tests/test_gpu_fuzz.py -m gpu -x -q --timeout=120 --timeout-method=thread 2>&1 | tail -2
timeout 200 python tools/perf_probe.py 2>&1 | grep -E "rep 2|timeline"
