#!/bin/bash
# round 2, visit M (8 GPUs): multi-GPU tests, then scaling with the cost-model dispatch and interleaved sector shards
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_pipeline.py tests/test_gpu_tile.py -x -q --timeout=600 -k "multirank or devices or two_gpus or tile" > gpurun_out/r02m_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02m_tests.log; tail -4 gpurun_out/r02m_tests.log
source tools/gpu_r02l.sh.lib
for n in 1 2 4 8; do run scale120k $n; done
for n in 1 2 4 8; do run scale500k $n --beams 128 --az 3907; done
