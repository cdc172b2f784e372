#!/bin/bash
# round 2, visit F: the whole GPU suite on the current code, sort cost (launch list), bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r02f_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02f_pytest.log
tail -12 gpurun_out/r02f_pytest.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'tile_|RadixSort|icp_init|nn_tile' -c 60 --csv --log-file gpurun_out/r02f_launches.csv python tools/ncu_target.py 5000000 4 > gpurun_out/r02f_launches.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open('gpurun_out/r02f_launches.csv')) if len(r) > 10]
hdr = rows[0]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value")
for r in rows[-26:]:
    print(r[kn][:90].ljust(92), r[mv])
PY
timeout 900 python bench.py --steps 20 > gpurun_out/r02f_bench.json 2> gpurun_out/r02f_bench.err; echo "bench rc=$?"; cat gpurun_out/r02f_bench.json; tail -5 gpurun_out/r02f_bench.err
