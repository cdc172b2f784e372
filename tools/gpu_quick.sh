#!/bin/bash
# quick GPU visit: parity tests + smoke + bench + stream (no ncu)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=400 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 400 python bench.py $BENCH_ARGS > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['avg_launch_us'], d['roofline']['traffic'], 'cpu', d['cpu_baseline']['value'], d['clocks'])"
tail -3 gpurun_out/bench.err
timeout 600 python tools/stream_bench.py --frames 300 --cpu-frames 20 > gpurun_out/stream.json 2> gpurun_out/stream.err; echo rc=$?; cat gpurun_out/stream.json; tail -3 gpurun_out/stream.err
