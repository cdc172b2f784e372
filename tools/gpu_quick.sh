#!/bin/bash
# quick GPU visit: parity tests + smoke + bench (no ncu)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=150 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python bench.py $BENCH_ARGS > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err
