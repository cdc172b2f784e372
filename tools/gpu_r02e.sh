#!/bin/bash
# round 2, visit E: register budgets / staging sizes, tail timeline, sort cost (launch list), full-size parity, bench with dispatch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tile.py tests/test_gpu_fullsize.py -x -q --timeout=400 > gpurun_out/r02e_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02e_tests.log
tail -6 gpurun_out/r02e_tests.log
TILE_TIMELINE=1 timeout 900 python tools/tile_probe.py legacy,tile,tile_minb4,tile_minb4_stage2560,tile_minb6_stage1408,tile_minb8_stage704 2000,15000,60000,120000 > gpurun_out/r02e_tile_probe.jsonl 2> gpurun_out/r02e_tile_probe.err; echo "rc=$?"
cut -c1-200 gpurun_out/r02e_tile_probe.jsonl; tail -5 gpurun_out/r02e_tile_probe.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/r02e_launches.csv python tools/ncu_target.py 5000000 4 > gpurun_out/r02e_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r02e_launches.csv')) if len(r) > 10]
hdr = rows[0]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value")
for r in rows[1:61]:
    print(r[kn][:70].ljust(72), r[mv])
PY
timeout 900 python bench.py --steps 20 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err; echo "bench rc=$?"; cat gpurun_out/r02e_bench.json; tail -5 gpurun_out/r02e_bench.err
