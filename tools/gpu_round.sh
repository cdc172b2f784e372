#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench, launch list, one full ncu capture of the search kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; lscpu | grep 'Model name' >> gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -q --timeout=400 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nn_search -s 40 -c 2 -f -o gpurun_out/prof_nn \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu_full.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench.json; tail -3 gpurun_out/bench.err; cat gpurun_out/bench_reference.json
timeout 150 python tools/stream_bench.py --frames 300 --cpu-frames 40 > gpurun_out/stream_300.json 2> gpurun_out/stream_300.err; tail -c 900 gpurun_out/stream_300.json
