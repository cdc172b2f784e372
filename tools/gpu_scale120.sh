#!/bin/bash
# driver-style scaling run on the bench config (120 k-point scan), fused peer exchange
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for n in 2 4 8; do
  if [ $n -gt $NG ]; then break; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale120k_n$n.json 2> gpurun_out/scale120k_n$n.err
  echo "N=$n rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/scale120k_n$n.json')); print(d['n_gpus'], d['value'], d['e2e']['value'], d['roofline']['avg_launch_us'], d['gpu_launches'])" || tail -5 gpurun_out/scale120k_n$n.err
done
