#!/bin/bash
# round 2, visit AH: final check of the committed code — whole GPU suite, smoke, bench (both arms), launch list + full ncu capture,
# the 1000-frame streaming drive (configs[2]) and the map-size sweep (configs[4])
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r02ah_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 > gpurun_out/r02ah_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02ah_pytest.log
tail -6 gpurun_out/r02ah_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02ah_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02ah_smoke.log; tail -2 gpurun_out/r02ah_smoke.log
timeout 900 python bench.py > gpurun_out/r02ah_bench.json 2> gpurun_out/r02ah_bench.err; echo "bench rc=$?"; cat gpurun_out/r02ah_bench.json; tail -3 gpurun_out/r02ah_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02ah_bench_reference.json 2> gpurun_out/r02ah_bench_reference.err; cut -c1-300 gpurun_out/r02ah_bench_reference.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 400 --csv --log-file gpurun_out/r02ah_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pipeline --no-hbm-regime > gpurun_out/r02ah_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:nn_tile -s 4 -c 1 -f -o gpurun_out/r02ah_tile \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-pipeline --no-hbm-regime > gpurun_out/r02ah_bench_under_ncu_full.log 2>&1; echo "ncu rc=$?"
timeout 600 python tools/stream_bench.py --frames 1000 --cpu-frames 60 > gpurun_out/r02ah_stream_1000frames.json 2> gpurun_out/r02ah_stream_1000frames.err; echo "stream rc=$?"; cut -c1-900 gpurun_out/r02ah_stream_1000frames.json
timeout 600 python tools/map_sweep.py --sizes 1,5,20,50 > gpurun_out/r02ah_map_sweep.jsonl 2> gpurun_out/r02ah_map_sweep.err; echo "sweep rc=$?"; cut -c1-260 gpurun_out/r02ah_map_sweep.jsonl
