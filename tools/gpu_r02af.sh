#!/bin/bash
# visit AF (2 GPUs): the multi-rank tests (two processes, one process / two GPUs) and the bench at N = 2 on the current code
mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_multirank.py tests/test_gpu_pipeline.py -x -q -m gpu --timeout 200 --timeout-method=thread > gpurun_out/r02af_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r02af_pytest.log; tail -3 gpurun_out/r02af_pytest.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02af_bench_n2.json 2> gpurun_out/r02af_bench_n2.err; echo "n2 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02af_bench_n2.json')); print('N=2', d['value'], d['e2e']['value'], d['roofline']['us_per_iteration'], d.get('sharded_pose_delta_m'))" || tail -5 gpurun_out/r02af_bench_n2.err
timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-pipeline --no-hbm-regime > gpurun_out/r02af_bench_n1.json 2> gpurun_out/r02af_bench_n1.err
python -c "
import json; d=json.load(open('gpurun_out/r02af_bench_n1.json')); print('N=1', d['value'], d['e2e']['value'], d['roofline']['us_per_iteration'])"
