#!/bin/bash
# round 2, visit Z: L1 prefetch hints in the tile kernel (A/B against the build of commit bb1cf9d), every-block-steps threshold of the tile search
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tile.py -x -q -m gpu --timeout 150 --timeout-method=thread > gpurun_out/r02z_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02z_pytest.log; tail -3 gpurun_out/r02z_pytest.log
B="--steps 20 --no-cpu-baseline --no-pipeline --no-hbm-regime"
for v in prev new prev2 new2; do
  if [ ${v:0:4} = prev ]; then export SAGE_ICP_LIB=$PWD/build/variants/libsage_prev.so; else unset SAGE_ICP_LIB; fi
  timeout 300 python bench.py $B > gpurun_out/r02z_bench_$v.json 2> gpurun_out/r02z_bench_$v.err; echo "$v rc=$?"
done
unset SAGE_ICP_LIB
python - <<'PY'
import json
for f in ("prev", "new", "prev2", "new2"):
    try:
        d = json.load(open(f"gpurun_out/r02z_bench_{f}.json")); r = d["roofline"]
        print(f, "scans/s", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"], 1), "us/iter", round(r["us_per_iteration"], 2))
    except Exception as e:
        print(f, "failed", e)
PY
TILE_TIMELINE=0 timeout 300 python tools/tile_probe.py tile_elected,tile_every_block 15000,30000,60000,120000 > gpurun_out/r02z_tile_probe.jsonl 2> gpurun_out/r02z_tile_probe.err; echo "probe rc=$?"
cut -c1-110 gpurun_out/r02z_tile_probe.jsonl
