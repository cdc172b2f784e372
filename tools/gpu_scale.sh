#!/bin/bash
# BASELINE configs[3]: 500 k-point dense scan (128 x 3907 rays), query shard + NCCL all-reduce of the normal equations, 1/2/4/8 GPUs
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for n in 1 2 4 8; do
  if [ $n -gt $NG ]; then break; fi
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --beams 128 --az 3907 --no-cpu-baseline > gpurun_out/scale500k_n1.json 2> gpurun_out/scale500k_n1.err
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 20 --warmup 3 --beams 128 --az 3907 --no-cpu-baseline > gpurun_out/scale500k_n$n.json 2> gpurun_out/scale500k_n$n.err
  fi
  echo "N=$n rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/scale500k_n$n.json")); print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["e2e"]["value"], d["roofline"]["avg_launch_us"])
except Exception as e: print("no json", e)
PY
done
