#!/bin/bash
# visit AB: late draw of the next unit — tile_probe at shard sizes, A/B against the build of commit bb1cf9d, then the 15 k timeline
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tile.py -x -q -m gpu --timeout 150 --timeout-method=thread 2>&1 | tail -2
for v in prev new prev2 new2; do
  if [ ${v:0:4} = prev ]; then export SAGE_ICP_LIB=$PWD/build/variants/libsage_prev.so; else unset SAGE_ICP_LIB; fi
  echo "== $v"; TILE_TIMELINE=0 timeout 300 python tools/tile_probe.py tile 8000,15000,30000,60000,120000 2>/dev/null | cut -c1-70
done
unset SAGE_ICP_LIB
TILE_TIMELINE_N=15000 timeout 300 python tools/tile_probe.py tile 15000 2>&1 | cut -c1-200 | tail -20
