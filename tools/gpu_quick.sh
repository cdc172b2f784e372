#!/bin/bash
mkdir -p gpurun_out
timeout 100 python __graft_entry__.py smoke 2>&1 | tail -4 | tee gpurun_out/smoke.log
