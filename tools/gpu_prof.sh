#!/bin/bash
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:nn_search -s 4 -c 1 -f -o gpurun_out/prof_nn \
    python tools/perf_probe.py > gpurun_out/ncu_full.log 2>&1; tail -2 gpurun_out/ncu_full.log
