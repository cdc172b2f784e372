#!/bin/bash
# scratch: streaming timing with / without the persistent GN-loop kernel, alternating, slowest frames listed
mkdir -p gpurun_out
for r in 1 2; do
SAGE_PERSISTENT_MAX=0 timeout 100 python tools/stream_bench.py --frames 200 --cpu-frames 6 > gpurun_out/stream_per_launch_$r.json 2> gpurun_out/stream_per_launch.err
timeout 100 python tools/stream_bench.py --frames 200 --cpu-frames 6 > gpurun_out/stream_persistent_$r.json 2> gpurun_out/stream_persistent.err
done
python - <<'PY'
import json
for f in ("per_launch_1","persistent_1","per_launch_2","persistent_2"):
    d=json.load(open(f"gpurun_out/stream_{f}.json"))
    print(f, {k:d[k] for k in ("gpu_frames_per_s","gpu_ms_per_frame_median","gpu_ms_per_frame_p99","mean_t_icp_ms","slowest_frames (index, ms)")})
PY
