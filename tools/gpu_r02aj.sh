#!/bin/bash
# visit AJ (8 GPUs): the bench configuration at N = 8 on the final code (driver-style launch)
mkdir -p gpurun_out
n=8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29708 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r02aj_scale120k_n$n.json 2> gpurun_out/r02aj_scale120k_n$n.err
echo "N=$n rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02aj_scale120k_n$n.json')); print(d['n_gpus'], d['value'], d['e2e']['value'], d['roofline']['us_per_iteration'], d.get('sharded_pose_delta_m'))" || tail -5 gpurun_out/r02aj_scale120k_n$n.err
