#!/bin/bash
# round 2, visit L (8 GPUs): scaling of the bench config (120 k-point scan) and of BASELINE configs[3] (500 k-point scan), fused peer exchange
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "GPUs: $NG"
run() {  # name n extra-args
  local name=$1 n=$2; shift 2
  if [ $n -eq 1 ]; then
    timeout 400 python bench.py --gpus 1 --steps 12 --warmup 3 --no-cpu-baseline --no-pipeline --no-hbm-regime "$@" > gpurun_out/r02l_${name}_n1.json 2> gpurun_out/r02l_${name}_n1.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 12 --warmup 3 --no-cpu-baseline --no-pipeline --no-hbm-regime "$@" > gpurun_out/r02l_${name}_n$n.json 2> gpurun_out/r02l_${name}_n$n.err
  fi
  echo "$name N=$n rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r02l_${name}_n$n.json")); print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, "e2e", d["e2e"]["value"], "us/iter", d["roofline"]["us_per_iteration"], "delta", d.get("sharded_pose_delta_m"))
except Exception as e: print("no json", e)
PY
  tail -2 gpurun_out/r02l_${name}_n$n.err | cut -c1-300
}
for n in 1 2 4 8; do [ $n -le $NG ] && run scale120k $n; done
for n in 1 2 4 8; do [ $n -le $NG ] && run scale500k $n --beams 128 --az 3907; done
