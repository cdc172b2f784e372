#!/usr/bin/env python
"""BASELINE configs[2]: 1x B200 streaming — an N-frame synthetic KITTI-shaped drive (64 x 1875 rays, 1 m/frame) through the
full sageICP::RegisterFrame with the map updated incrementally on the device.  Prints one JSON line: GPU frames/s
(wall clock around the C-ABI call with pinned host buffers, map update included), the oracle's frames/s on a bounded
prefix of the same drive, per-frame pose parity over that prefix.  Not the contract bench (bench.py is)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=1000)
    ap.add_argument("--cpu-frames", type=int, default=40)
    ap.add_argument("--beams", type=int, default=64)
    ap.add_argument("--az", type=int, default=1875)
    ap.add_argument("--dynamic-filter", action="store_true", help="dynamic_vehicle_filter = true (ros/launch/odometry.launch.py:50)")
    ap.add_argument("--pageable", action="store_true", help="hand the scans over in pageable memory, as the reference's node does (a std::vector)")
    a = ap.parse_args()
    import torch
    import sage_icp_b200 as sg
    from sage_icp_b200 import synthetic as syn
    from oracle import oracle_py as orc
    cfg = sg.launch_config(dynamic_vehicle_filter=a.dynamic_filter)
    traj = syn.trajectory(a.frames)
    gp = sg.SagePipeline(cfg)
    op = orc.OraclePipeline(cfg, threads=orc.max_threads(), evict_faithful=False)
    pin = torch.empty((a.beams * a.az, 4), dtype=torch.float64).pin_memory()
    t_gpu, t_cpu, iters, nsrc, nds, worst = [], [], [], [], [], [0.0, 0.0]
    t_icps, t_alls = [], []
    l0 = sg.launch_count()
    for i in range(a.frames):
        scan = syn.make_scan(i, tuple(traj[i]), n_beams=a.beams, n_az=a.az)
        pin.copy_(torch.from_numpy(scan))
        buf = np.ascontiguousarray(scan) if a.pageable else pin.numpy()
        torch.cuda.synchronize()
        t = time.perf_counter()
        pg, t_icp, t_all = gp.register_frame(buf)
        gp.map().num_voxels()  # waits for the asynchronous map update: the frame is fully done
        t_gpu.append(time.perf_counter() - t)
        iters.append(gp.last_iterations()); nsrc.append(len(gp.last_source())); t_icps.append(t_icp); t_alls.append(t_all)
        if i < a.cpu_frames:
            t = time.perf_counter()
            po, _, _ = op.register_frame(scan)
            t_cpu.append(time.perf_counter() - t)
            worst[0] = max(worst[0], float(np.linalg.norm(pg[:3] - po[:3])))
            worst[1] = max(worst[1], 2 * float(min(np.linalg.norm(pg[3:] - po[3:]), np.linalg.norm(pg[3:] + po[3:]))))
    warm = min(5, a.frames // 2)
    g, c = np.array(t_gpu[warm:]), np.array(t_cpu[warm:])
    print(json.dumps({
        "workload": "BASELINE configs[2]: streaming RegisterFrame, synthetic KITTI-shaped drive, incremental Update on device",
        "frames": a.frames, "host_buffers": "pageable" if a.pageable else "pinned", "dynamic_vehicle_filter": bool(a.dynamic_filter), "rays_per_scan": a.beams * a.az, "gpu_frames_per_s": float(1.0 / g.mean()), "gpu_ms_per_frame_median": float(1e3 * np.median(g)),
        "gpu_ms_per_frame_p99": float(1e3 * np.percentile(g, 99)),
        "slowest_frames (index, ms)": [(int(i) + warm, round(float(1e3 * g[i]), 3)) for i in np.argsort(g)[-5:][::-1]], "mean_gn_iterations": float(np.mean(iters)),
        "mean_t_icp_ms": float(1e3 * np.mean(t_icps[warm:])), "mean_t_all_ms (front end + icp, reference meaning)": float(1e3 * np.mean(t_alls[warm:])), "mean_queries": float(np.mean(nsrc)),
        "map_voxels_end": gp.map().num_voxels(), "map_points_end": gp.map().num_points(), "gpu_launches_per_frame": (sg.launch_count() - l0) / a.frames,
        "cpu_port_frames_per_s": float(1.0 / c.mean()), "cpu_cores": orc.max_threads(), "cpu_frames_timed": int(len(c)),
        "parity_prefix_frames": a.cpu_frames, "max_pose_delta_m": worst[0], "max_pose_delta_rad": worst[1],
        "distance_driven_m": float(traj[-1][0]), "final_x_estimate_m": float(gp.poses()[-1][0])}))


if __name__ == "__main__":
    main()
