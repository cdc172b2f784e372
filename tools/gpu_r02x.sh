#!/bin/bash
# round 2, visit X: every block steps for small scans (default) vs the elected block; per-query kernel compile variants in the DRAM-bound regime
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_streaming.py tests/test_gpu_core.py -x -q -m gpu \
    --timeout 150 --timeout-method=thread > gpurun_out/r02x_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02x_pytest.log; tail -4 gpurun_out/r02x_pytest.log
SMALL_PROBE_VARIANTS="default,elected" timeout 300 python tools/small_probe.py 700,2100,5000,12000 > gpurun_out/r02x_small_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r02x_small_probe.log
for v in default u8m4 u4m3 u8m3 u8m2; do
  if [ $v = default ]; then unset SAGE_ICP_LIB; else export SAGE_ICP_LIB=$PWD/build/variants/libsage_$v.so; fi
  timeout 200 python tools/hbm_target.py 20000000 4 > gpurun_out/r02x_hbm_$v.json 2> gpurun_out/r02x_hbm_$v.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/r02x_hbm_$v.json')); print('$v', 'us/iter', round(d['us_per_iteration'],1), 'frac', round(d['frac'],3))" || tail -3 gpurun_out/r02x_hbm_$v.err
done
