#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_core.py tests/test_gpu_search_exactness.py -m gpu -x -q --timeout=60 --timeout-method=thread 2>&1 | tail -2
echo "== default"; timeout 200 python tools/perf_probe.py 2>&1 | grep -E "rep 2|timeline|phase"
for v in t256_u2 t256_u6; do echo "== variant $v"; SAGE_ICP_LIB=$PWD/sage_icp_b200/lib/variant_$v.so timeout 200 python tools/perf_probe.py 2>&1 | grep -E "rep 2|timeline|home scanned|neighbours"; done
