#!/bin/bash
# round 2, visit R: every-block-steps persistent loops — parity (whole GPU suite), A/B timing, streaming
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 > gpurun_out/r02r_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02r_pytest.log; tail -6 gpurun_out/r02r_pytest.log
TILE_TIMELINE=0 timeout 900 python tools/tile_probe.py legacy,tile 700,2000,8000,15000,120000 2> gpurun_out/r02r_probe.err | cut -c1-110
SAGE_STEP_EVERYWHERE=0 TILE_TIMELINE=0 timeout 900 python tools/tile_probe.py legacy,tile 700,2000,8000,15000,120000 2>> gpurun_out/r02r_probe.err | cut -c1-110
timeout 600 python tools/stream_bench.py --frames 400 --cpu-frames 5 > gpurun_out/r02r_stream_400.json 2> gpurun_out/r02r_stream_400.err; python -c "
import json; d=json.load(open('gpurun_out/r02r_stream_400.json')); print({k: d[k] for k in ('gpu_frames_per_s','gpu_ms_per_frame_median','gpu_ms_per_frame_p99','mean_t_icp_ms','mean_gn_iterations','max_pose_delta_m')})"
SAGE_STEP_EVERYWHERE=0 timeout 600 python tools/stream_bench.py --frames 400 --cpu-frames 5 > gpurun_out/r02r_stream_400_old.json 2> gpurun_out/r02r_stream_400_old.err; python -c "
import json; d=json.load(open('gpurun_out/r02r_stream_400_old.json')); print({k: d[k] for k in ('gpu_frames_per_s','gpu_ms_per_frame_median','gpu_ms_per_frame_p99','mean_t_icp_ms','mean_gn_iterations','max_pose_delta_m')})"
