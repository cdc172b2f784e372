#!/bin/bash
# round 2, visit N (1 GPU): configs[2] streaming drive, configs[4] map sweep, DRAM traffic of the uniform-query regime
mkdir -p gpurun_out
timeout 900 python tools/stream_bench.py --frames 1000 --cpu-frames 60 > gpurun_out/r02n_stream_1000.json 2> gpurun_out/r02n_stream_1000.err; echo "stream rc=$?"; tail -c 1500 gpurun_out/r02n_stream_1000.json
timeout 900 python tools/map_sweep.py --sizes 1,5,20,50 > gpurun_out/r02n_map_sweep.jsonl 2> gpurun_out/r02n_map_sweep.err; echo "sweep rc=$?"; cut -c1-330 gpurun_out/r02n_map_sweep.jsonl
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:nn_search_kernel -s 30 -c 20 --csv --log-file gpurun_out/r02n_hbm_regime_ncu.csv python tools/hbm_target.py 50000000 4 > gpurun_out/r02n_hbm_target.log 2>&1; echo "ncu rc=$?"; tail -1 gpurun_out/r02n_hbm_target.log | cut -c1-600
tail -12 gpurun_out/r02n_hbm_regime_ncu.csv | cut -c1-200
