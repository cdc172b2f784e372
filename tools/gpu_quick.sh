#!/bin/bash
timeout 300 python tools/sort_probe.py 2>&1 | tail -6
