#!/bin/bash
# visit AI (4 GPUs): the bench configuration at N = 4 on the final code (driver-style launch), then N = 1 on the same box
mkdir -p gpurun_out
n=4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29704 bench.py --gpus $n --steps 20 --warmup 3 > gpurun_out/r02ai_scale120k_n$n.json 2> gpurun_out/r02ai_scale120k_n$n.err
echo "N=$n rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02ai_scale120k_n$n.json')); print(d['n_gpus'], d['value'], d['e2e']['value'], d['roofline']['us_per_iteration'], d.get('sharded_pose_delta_m'))" || tail -5 gpurun_out/r02ai_scale120k_n$n.err
timeout 200 python bench.py --steps 20 --no-cpu-baseline --no-pipeline --no-hbm-regime > gpurun_out/r02ai_scale120k_n1.json 2> gpurun_out/r02ai_scale120k_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02ai_scale120k_n1.json')); print(1, d['value'], d['e2e']['value'], d['roofline']['us_per_iteration'])"
