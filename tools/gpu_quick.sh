#!/bin/bash
echo "== thread mode forced"; SAGE_NO_ALL_WARP=1 timeout 300 python tools/mode_probe.py 2>&1 | tail -7
echo "== warp mode forced"; SAGE_ALL_WARP_MAX=100000000 timeout 300 python tools/mode_probe.py 2>&1 | tail -7
