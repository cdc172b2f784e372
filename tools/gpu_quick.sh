#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_core.py tests/test_gpu_pipeline.py -m gpu -x -q --timeout=120 --timeout-method=thread 2>&1 | tail -2
echo "== default (256 threads)"; timeout 200 python tools/perf_probe.py 2>&1 | grep -E "rep 2|timeline"
for v in t128 t512; do echo "== variant $v"; SAGE_ICP_LIB=$PWD/sage_icp_b200/lib/variant_$v.so timeout 200 python tools/perf_probe.py 2>&1 | grep -E "rep 2|timeline"; done
