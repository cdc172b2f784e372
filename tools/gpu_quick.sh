#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=400 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
timeout 200 python tools/perf_probe.py 2>&1 | grep -E "rep 2|timeline"
timeout 500 python tools/stream_bench.py --frames 300 --cpu-frames 300 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('gpu_frames_per_s','gpu_ms_per_frame_median','mean_t_icp_ms','max_pose_delta_m','max_pose_delta_rad','mean_gn_iterations')})"
