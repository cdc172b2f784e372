#!/bin/bash
# visit AG: the second Voxelize pass in one single-block kernel — front-end parity tests, then the stream with and without it
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_streaming.py tests/test_gpu_vs_reference_build.py tests/test_golden.py -x -q -m gpu --timeout 200 --timeout-method=thread > gpurun_out/r02ag_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02ag_pytest.log; tail -3 gpurun_out/r02ag_pytest.log
for v in small general small2 general2; do
  if [ ${v:0:7} = general ]; then export SAGE_FE_SMALL=0; else unset SAGE_FE_SMALL; fi
  SAGE_FE_TRACE=1 timeout 200 python tools/stream_bench.py --frames 300 --cpu-frames 0 > gpurun_out/r02ag_stream_$v.json 2> gpurun_out/r02ag_stream_$v.err
  grep "sage front end" gpurun_out/r02ag_stream_$v.err | cut -c1-330
  python -c "
import json; d=json.load(open('gpurun_out/r02ag_stream_$v.json')); print('$v', 'median ms', round(d['gpu_ms_per_frame_median'],3), 'p99', round(d['gpu_ms_per_frame_p99'],3), 't_icp', round(d['mean_t_icp_ms'],3), 't_all', round(d['mean_t_all_ms (front end + icp, reference meaning)'],3), 'launches/frame', d['gpu_launches_per_frame'], d['slowest_frames (index, ms)'][:3])"
done
