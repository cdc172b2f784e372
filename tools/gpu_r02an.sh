#!/bin/bash
# visit AN: last look at the library as committed — smoke, a short bench line, the core parity tests
mkdir -p gpurun_out
timeout 60 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 60 python bench.py --steps 20 --no-cpu-baseline --no-pipeline --no-hbm-regime > gpurun_out/r02an_bench.json 2> gpurun_out/r02an_bench.err; python -c "
import json; d=json.load(open('gpurun_out/r02an_bench.json')); print('bench', round(d['value'],1), round(d['e2e']['value'],1), round(d['roofline']['us_per_iteration'],2), d['gpu_launches'], d['clocks'])"
timeout 50 python -m pytest tests/test_gpu_core.py -x -q -m gpu --timeout 40 --timeout-method=thread 2>&1 | tail -2
