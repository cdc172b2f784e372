#!/bin/bash
mkdir -p gpurun_out
python - <<'PY'
import re
s=open('tools/tile_probe.py').read()
s=s.replace('"tile_minb4": {"SAGE_TILE_MIN": "1", "SAGE_TILE_MINB": "4"},','"tile_minb4": {"SAGE_TILE_MIN": "1", "SAGE_TILE_MINB": "4"},\n    "tile_minb5": {"SAGE_TILE_MIN": "1", "SAGE_TILE_MINB": "5"},\n    "tile_minb5_se2": {"SAGE_TILE_MIN": "1", "SAGE_TILE_MINB": "5", "SAGE_STEP_EVERYWHERE": "2"},\n    "tile_minb6_1408": {"SAGE_TILE_MIN": "1", "SAGE_TILE_MINB": "6", "SAGE_TILE_STAGE": "1408"},')
s=s.replace('KEYS = ("SAGE_TILE",','KEYS = ("SAGE_STEP_EVERYWHERE", "SAGE_TILE",')
open('tools/tile_probe.py','w').write(s)
PY
TILE_TIMELINE=1 timeout 900 python tools/tile_probe.py legacy,tile,tile_minb5,tile_minb5_se2,tile_minb6_1408 15000,60000,120000 5000000 SAGE_TILE_MINB=5 > gpurun_out/r02t_probe.jsonl 2> gpurun_out/r02t_probe.err
cut -c1-110 gpurun_out/r02t_probe.jsonl; tail -3 gpurun_out/r02t_probe.err
