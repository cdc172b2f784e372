#!/bin/bash
# round 2, visit H (2 GPUs): heavy-first scheduling, one handle on two GPUs, the two-process sharded tests, bench at N = 2
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02h_gpus.txt
timeout 900 python -m pytest tests/test_gpu_tile.py tests/test_gpu_multirank.py tests/test_gpu_pipeline.py -x -q --timeout=600 -k "tile or multirank or devices or two_gpus" > gpurun_out/r02h_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02h_tests.log
tail -8 gpurun_out/r02h_tests.log
TILE_TIMELINE=1 timeout 900 python tools/tile_probe.py legacy,tile 2000,15000,60000,120000 > gpurun_out/r02h_tile_probe.jsonl 2> gpurun_out/r02h_tile_probe.err; echo "rc=$?"
cut -c1-200 gpurun_out/r02h_tile_probe.jsonl; tail -5 gpurun_out/r02h_tile_probe.err
timeout 900 python bench.py --steps 20 --no-hbm-regime --no-pipeline > gpurun_out/r02h_bench_n1.json 2> gpurun_out/r02h_bench_n1.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r02h_bench_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02h_bench_n2.json 2> gpurun_out/r02h_bench_n2.err; echo "bench2 rc=$?"; cut -c1-800 gpurun_out/r02h_bench_n2.json; tail -5 gpurun_out/r02h_bench_n2.err
