#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tile.py tests/test_gpu_keyframe.py -x -q --timeout=300 > gpurun_out/r02g_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02g_tests.log
tail -5 gpurun_out/r02g_tests.log
TILE_TIMELINE=1 timeout 900 python tools/tile_probe.py legacy,tile,tile_launch_per_iter 2000,8000,15000,30000,60000,120000 > gpurun_out/r02g_tile_probe.jsonl 2> gpurun_out/r02g_tile_probe.err; echo "rc=$?"
cut -c1-200 gpurun_out/r02g_tile_probe.jsonl; tail -5 gpurun_out/r02g_tile_probe.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tile_|RadixSort|icp_init|nn_tile' -c 40 --csv --log-file gpurun_out/r02g_launches.csv python tools/ncu_target.py 5000000 3 > gpurun_out/r02g_launches.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r02g_launches.csv')) if len(r) > 10]
hdr = rows[0]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value")
for r in rows[-11:]:
    print(r[kn][:90].ljust(92), r[mv])
PY
timeout 900 python bench.py --steps 20 > gpurun_out/r02g_bench.json 2> gpurun_out/r02g_bench.err; echo "bench rc=$?"; cut -c1-1500 gpurun_out/r02g_bench.json; tail -5 gpurun_out/r02g_bench.err
