"""Scratch: the DRAM-bound regime of bench.py (roofline_hbm_regime) on its own, for ncu to attach to."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sage_icp_b200 as sg
import bench
n_map = int(sys.argv[1]) if len(sys.argv) > 1 else 50_000_000
base = bench.make_map_points(5_000_000)
peak, _ = bench.measured_peak_gbs()
out = bench.hbm_regime_leg(sg, torch, 0, n_map, base, bench.street_half_length(5_000_000), peak, reps=int(sys.argv[2]) if len(sys.argv) > 2 else 6)
print(json.dumps(out))
