#!/bin/bash
mkdir -p gpurun_out
for comm in peer nccl; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 --comm $comm --no-cpu-baseline > gpurun_out/bench_n2_$comm.json 2> gpurun_out/bench_n2_$comm.err; echo "$comm rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/bench_n2_$comm.json')); print('$comm', d['value'], d['e2e']['value'], d['roofline']['avg_launch_us'], d['gpu_launches'])" || tail -5 gpurun_out/bench_n2_$comm.err
done
