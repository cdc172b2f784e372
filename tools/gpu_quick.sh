#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=400 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
echo "== default policy"; timeout 300 python tools/mode_probe.py 2>&1 | tail -7 | tr '\n' ' '
