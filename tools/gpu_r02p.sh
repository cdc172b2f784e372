#!/bin/bash
# round 2, visit P (1 GPU): L2 prefetch in the per-query kernel (uniform regime + small scans), growth trace of a streaming drive, new tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_core.py tests/test_gpu_search_exactness.py -x -q --timeout=300 2>&1 | tail -3
timeout 600 python tools/tile_probe.py legacy,tile 2000,8000,15000,120000 2> gpurun_out/r02p_probe.err | cut -c1-110
TILE_TIMELINE=0 timeout 900 python tools/hbm_target.py 50000000 6 2> gpurun_out/r02p_hbm.err | cut -c1-700
SAGE_TRACE_GROWTH=1 timeout 600 python tools/stream_bench.py --frames 700 --cpu-frames 5 > gpurun_out/r02p_stream_700.json 2> gpurun_out/r02p_stream_700.err; grep "\[sage\]" gpurun_out/r02p_stream_700.err | head -20; python -c "
import json; d=json.load(open('gpurun_out/r02p_stream_700.json')); print({k: d[k] for k in ('gpu_frames_per_s','gpu_ms_per_frame_median','gpu_ms_per_frame_p99','slowest_frames (index, ms)')})"
