#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --timeout=200 --timeout-method=thread 2>&1 | tail -4
timeout 600 python tools/stream_bench.py --frames 200 --cpu-frames 6 --dynamic-filter > gpurun_out/stream_dyn.json 2> gpurun_out/stream_dyn.err; echo rc=$?; cat gpurun_out/stream_dyn.json; tail -3 gpurun_out/stream_dyn.err
