#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout=400 --timeout-method=thread > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['avg_launch_us'], d['roofline']['traffic'], 'cpu', d['cpu_baseline']['value'], d['clocks'])"
timeout 600 python tools/stream_bench.py --frames 300 --cpu-frames 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('gpu_frames_per_s','gpu_ms_per_frame_median','mean_t_icp_ms','max_pose_delta_m')})"
