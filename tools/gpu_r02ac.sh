#!/bin/bash
# visit AC: 120 k timeline with the late draw
mkdir -p gpurun_out
timeout 300 python tools/tile_probe.py tile 120000 2>&1 | cut -c1-200 | tail -21
