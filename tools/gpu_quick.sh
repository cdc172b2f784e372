#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=400 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python tools/stream_bench.py --frames 300 --cpu-frames 20 > gpurun_out/stream.json 2> gpurun_out/stream.err; echo rc=$?; cat gpurun_out/stream.json; tail -3 gpurun_out/stream.err
