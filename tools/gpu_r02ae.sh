#!/bin/bash
# visit AE: which call stalls frame 27 of a fresh drive (150-400 ms in visit AD)?
mkdir -p gpurun_out
for k in 1 2; do
SAGE_TRACE_SLOW=3 SAGE_TRACE_GROWTH=1 timeout 200 python tools/stream_bench.py --frames 60 --cpu-frames 0 > gpurun_out/r02ae_stream_$k.json 2> gpurun_out/r02ae_stream_$k.err
grep "sage" gpurun_out/r02ae_stream_$k.err | cut -c1-250; python -c "
import json; d=json.load(open('gpurun_out/r02ae_stream_$k.json')); print(d['slowest_frames (index, ms)'], d['gpu_ms_per_frame_median'])"
done
