#!/bin/bash
# scratch: 2-GPU lock-step tests after the kernel-body refactor
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu --timeout 100 --timeout-method=thread 2>&1 | tail -6 | tee gpurun_out/quick_multirank.log
