#!/bin/bash
# scratch: validate the boundary changes (config deep copy, sage_set_devices) on the pipeline path
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pipeline.py tests/test_cpp_adaptor.py tests/test_capi_boundary.py -x -q --timeout 120 --timeout-method=thread 2>&1 | tail -6 | tee gpurun_out/quick_boundary.log
