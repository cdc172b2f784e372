#!/bin/bash
# round 2, visit W: pooled passes of tiny units (four to a pass, one warp each) — parity, then same-box A/B against the previous build
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_tile.py tests/test_gpu_search_exactness.py tests/test_gpu_core.py -x -q -m gpu \
    --timeout 150 --timeout-method=thread > gpurun_out/r02w_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02w_pytest.log; tail -4 gpurun_out/r02w_pytest.log
B="--steps 20 --no-cpu-baseline --no-pipeline --no-hbm-regime"
SAGE_ICP_LIB=$PWD/build/variants/libsage_prev.so timeout 300 python bench.py $B > gpurun_out/r02w_bench_prev.json 2> gpurun_out/r02w_bench_prev.err; echo "prev rc=$?"
SAGE_TILE_POOL=0 timeout 300 python bench.py $B > gpurun_out/r02w_bench_nopool.json 2> gpurun_out/r02w_bench_nopool.err; echo "nopool rc=$?"
timeout 300 python bench.py $B > gpurun_out/r02w_bench_pool.json 2> gpurun_out/r02w_bench_pool.err; echo "pool rc=$?"
python - <<'PY'
import json
for f in ("prev", "nopool", "pool"):
    try:
        d = json.load(open(f"gpurun_out/r02w_bench_{f}.json")); r = d["roofline"]
        print(f, "scans/s", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "us/iter", round(r["us_per_iteration"], 2), "launches", d["gpu_launches"])
    except Exception as e:
        print(f, "failed", e)
PY
SMALL_PROBE_VARIANTS="wide2" timeout 300 python tools/small_probe.py 700,2100,5000 > gpurun_out/r02w_small_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/r02w_small_probe.log
