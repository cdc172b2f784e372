#!/bin/bash
# visit AL: compile-time variants of the tile kernel's group size (two-level sum) and pair-list length, same box
mkdir -p gpurun_out
for v in base g32 g64 p512 base2 g32b; do
  case $v in base|base2) unset SAGE_ICP_LIB;; g32b) export SAGE_ICP_LIB=$PWD/build/variants/libsage_g32.so;; *) export SAGE_ICP_LIB=$PWD/build/variants/libsage_$v.so;; esac
  echo "== $v"; TILE_TIMELINE=0 timeout 200 python tools/tile_probe.py tile 15000,60000,120000 2>/dev/null | cut -c1-66
done
