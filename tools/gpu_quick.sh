#!/bin/bash
# scratch: new faithful-eviction tests first, then the map/pipeline regression subset
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_core.py::test_update_and_evict_faithful_mode tests/test_gpu_streaming.py::test_streaming_drive_with_the_reference_eviction_quirk -x -q -m gpu --timeout 200 --timeout-method=thread 2>&1 | tail -30 | tee gpurun_out/quick_faithful.log
timeout 300 python -m pytest tests/test_gpu_core.py tests/test_gpu_pipeline.py -x -q -m gpu --timeout 120 --timeout-method=thread 2>&1 | tail -8 | tee gpurun_out/quick_regress.log
