#!/bin/bash
# visit AM: group sums with all loads in flight — tile + exactness tests, then A/B against the build of commit f6e8ee1 (same box)
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_tile.py tests/test_gpu_search_exactness.py -x -q -m gpu --timeout 150 --timeout-method=thread 2>&1 | tail -2
for v in base new base2 new2; do
  case $v in base|base2) export SAGE_ICP_LIB=$PWD/build/variants/libsage_base.so;; *) unset SAGE_ICP_LIB;; esac
  echo "== $v"; TILE_TIMELINE=0 timeout 200 python tools/tile_probe.py tile 15000,120000 2>/dev/null | cut -c1-66
done
