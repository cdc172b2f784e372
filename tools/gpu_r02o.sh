#!/bin/bash
# round 2, visit O (8 GPUs): release/acquire exchange — multi-GPU tests, then the scaling points again
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_pipeline.py tests/test_gpu_core.py -x -q --timeout=600 -k "multirank or devices or two_gpus or rank_deficient" > gpurun_out/r02o_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02o_tests.log; tail -4 gpurun_out/r02o_tests.log
source tools/gpu_r02l.sh.lib
for n in 1 2 4 8; do run scale120k $n; done
for n in 1 8; do run scale500k $n --beams 128 --az 3907; done
