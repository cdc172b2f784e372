#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_core.py tests/test_gpu_search_exactness.py tests/test_golden.py -m gpu -x -q --timeout=60 --timeout-method=thread 2>&1 | tail -2
for lp in 8 5 3 2 1; do echo "== light probes $lp"; SAGE_LIGHT_PROBES=$lp timeout 200 python tools/perf_probe.py 2>&1 | grep -E "scan 0|rep 2|timeline"; done
echo "== local defer (old path)"; SAGE_LOCAL_DEFER=1 timeout 200 python tools/perf_probe.py 2>&1 | grep -E "rep 2|timeline"
