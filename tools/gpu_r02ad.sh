#!/bin/bash
# visit AD: pageable scans staged through pinned slots; where the front end's host time goes; pipeline leg of the bench
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_core.py tests/test_gpu_pipeline.py -x -q -m gpu --timeout 150 --timeout-method=thread 2>&1 | tail -3
SAGE_FE_TRACE=1 timeout 200 python tools/stream_bench.py --frames 300 --cpu-frames 0 > gpurun_out/r02ad_stream_pinned.json 2> gpurun_out/r02ad_stream_pinned.err; tail -2 gpurun_out/r02ad_stream_pinned.err | cut -c1-400
timeout 200 python tools/stream_bench.py --frames 300 --cpu-frames 0 --pageable > gpurun_out/r02ad_stream_pageable_staged.json 2> gpurun_out/r02ad_stream_pageable_staged.err
SAGE_HOST_STAGE=0 timeout 200 python tools/stream_bench.py --frames 300 --cpu-frames 0 --pageable > gpurun_out/r02ad_stream_pageable_plain.json 2> gpurun_out/r02ad_stream_pageable_plain.err
SAGE_HOST_STAGE_THREADS=8 timeout 200 python tools/stream_bench.py --frames 300 --cpu-frames 0 --pageable > gpurun_out/r02ad_stream_pageable_staged8.json 2> gpurun_out/r02ad_stream_pageable_staged8.err
python - <<'PY'
import json
for f in ("pinned", "pageable_staged", "pageable_plain", "pageable_staged8"):
    try:
        d = json.load(open(f"gpurun_out/r02ad_stream_{f}.json"))
        print(f, "frames/s", round(d["gpu_frames_per_s"], 1), "median ms", round(d["gpu_ms_per_frame_median"], 3), "p99", round(d["gpu_ms_per_frame_p99"], 3), "t_icp", round(d["mean_t_icp_ms"], 3),
              "t_all", round(d["mean_t_all_ms (front end + icp, reference meaning)"], 3), "iters", round(d["mean_gn_iterations"], 1), "queries", round(d["mean_queries"]))
    except Exception as e:
        print(f, "failed", e)
PY
timeout 300 python bench.py --steps 10 --no-cpu-baseline --no-hbm-regime > gpurun_out/r02ad_bench.json 2> gpurun_out/r02ad_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02ad_bench.json')); print('bench', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['us_per_iteration'],2)); print(json.dumps(d['pipeline'])[:900])"
