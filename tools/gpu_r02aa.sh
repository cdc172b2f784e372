#!/bin/bash
# visit AA: per-block timeline of the tile kernel on a 15 000-query scan (what a rank of an 8-GPU run works on)
mkdir -p gpurun_out
TILE_TIMELINE_N=15000 timeout 300 python tools/tile_probe.py tile 15000 > gpurun_out/r02aa_tile_probe_15k.log 2>&1; echo "rc=$?"; cut -c1-230 gpurun_out/r02aa_tile_probe_15k.log
