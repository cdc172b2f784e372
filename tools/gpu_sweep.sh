#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/map_sweep.py --sizes 1,5,20,50 > gpurun_out/map_sweep.jsonl 2> gpurun_out/map_sweep.err; echo rc=$?; cat gpurun_out/map_sweep.jsonl; tail -3 gpurun_out/map_sweep.err
timeout 300 python tools/stream_bench.py --frames 200 --cpu-frames 10 2>/dev/null | tail -1
