#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native SAGE-ICP registration hot path.

Metric (BASELINE.json): RegisterFrame scans/s on a 120 k-point labelled scan; NN-kernel GB/s vs the HBM roofline.
Workload at every N (BASELINE.json configs[1], kernel level): one synthetic 64x1875-ray labelled
scan fed directly as queries to sage_icp::RegisterFrame (core/Registration.cpp:113-141) against a pre-built 5 M-point
voxel map, exactly 10 Gauss-Newton iterations (threshold 0 => no early exit).  One "step" = one scan.

  value  : scans/s with the scan already resident in HBM (sage_core_register_frame_device)
  e2e    : scans/s through the C-ABI call with HOST buffers (pinned): H2D of the scan + D2H of the pose inside
  N > 1  : the scan's queries are sharded by index over the ranks (sage_shard_range), every rank holds a replica of
           the map, and the 17 normal-equation sums are all-reduced with NCCL every iteration  => "strong" scaling

--impl reference times the reference's CPU algorithm on ALL host cores (os.sched_getaffinity, whatever OMP_NUM_THREADS a
launcher set) on the same workload, every query of every scan: the oracle port (OpenMP) and, where its build travelled, the
reference's own sources compiled against stand-in third-party headers (oracle/_ref); the line's value is the faster of the two.

Extra objects of the main line (N = 1): `pipeline` = BASELINE configs[0], the function the north star names — one 120 k-point
scan through sage_register_frame (pageable host buffer) against a pre-built 1 M-point map, next to the CPU port single-threaded
("TBB off") and on all cores, pose delta; `roofline_hbm_regime` = the same kernel on queries spread uniformly over a 50 M-point
map (no reuse between queries: the DRAM-bound regime).  N > 1: `sharded_pose_delta_m` = |pose of the sharded run - pose of the
unsharded run on rank 0's GPU| (asserted <= 1e-10).
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MAX_DIST, KERNEL, SEM_TH, ITERS = 3.0, 1.0 / 3.0, 0.4, 10  # sigma = 1.0: 3*sigma, sigma/3 (pipeline/sageICP.cpp:80-85)
BASIC_LABELS = [40, 44, 48, 49, 50, 70, 72]
VOXEL_SIZE_MAP, BASIC, CRITICAL = 0.8, 20, 20  # ros/launch/odometry.launch.py:56-60


def street_half_length(n_map_points: int) -> float:
    # calibrated: 10 000 surface samples per metre of street leave ~4 600 map points per metre after the per-voxel
    # caps, ~18 points/voxel, ~250 voxels per metre (5 M points -> ~275 k voxels < 2^19)
    return max(120.0, n_map_points / 4600.0 / 2.0)


def make_map_points(n_map_points: int) -> np.ndarray:
    from sage_icp_b200 import synthetic as syn
    L = street_half_length(n_map_points)
    chunks, made, seed = [], 0, 1000
    want = int(2.17 * n_map_points)  # oversampling: the per-voxel caps drop about half of the samples
    while made < want:
        m = min(4_000_000, want - made)
        chunks.append(syn.sample_street_map(m, seed, -L, L))
        made += m
        seed += 1
    return np.concatenate(chunks)


def make_queries(step: int, n_beams: int, n_az: int, half_len: float):
    """Scan `step`, moved into the map frame with a perturbed pose (the ICP initial guess), SURVEY.md §8d."""
    from sage_icp_b200 import synthetic as syn
    span = max(0.0, half_len - 110.0)
    x = (-span + (2 * span) * ((step * 0.61803398875) % 1.0)) if span > 0 else 0.0
    scan = syn.make_scan(step, (x, 0.0, 0.0), n_beams=n_beams, n_az=n_az)
    guess = syn.pose7_from_xyyaw((x + 0.3, 0.1, math.radians(1.0)))
    return scan, guess


SECTORS_PER_RANK = 4


def shard_of(scan: np.ndarray, rank: int, world: int, n_az: int = 0) -> np.ndarray:
    """Rank `rank`'s share of a scan.  Any partition gives the same sums.  A spinning lidar delivers its points azimuth by azimuth,
    so a contiguous index range of a real scan is an angular sector; the synthetic scans are stored beam by beam (n_beams x n_az),
    so the same sector is a column range of every beam.  A sector is spatially compact — each rank works on its own cells of the
    map, which is what the tile search wants — and holds every beam, near and far.  Every rank gets SECTORS_PER_RANK thin sectors
    spread around the circle (sector s belongs to rank s % world), so that one rank does not get the whole dense side of the
    street: the ranks advance in lock-step, an iteration takes as long as the slowest of them.  Without the scan shape: chunks of
    32 consecutive points dealt round-robin."""
    if world == 1:
        return np.ascontiguousarray(scan)
    if n_az and len(scan) % n_az == 0:
        rows = scan.reshape(len(scan) // n_az, n_az, scan.shape[1])
        n_sec = world * SECTORS_PER_RANK
        cols = np.concatenate([np.arange(n_az * s // n_sec, n_az * (s + 1) // n_sec) for s in range(rank, n_sec, world)])
        return np.ascontiguousarray(rows[:, cols].reshape(-1, scan.shape[1]))
    n_chunks = (len(scan) + 31) // 32
    owner = np.repeat(np.arange(n_chunks) % world, 32)[: len(scan)]
    return np.ascontiguousarray(scan[owner == rank])


def algorithmic_bytes(n_q: int, occupied: int, candidates: int) -> float:
    """N_q*(16 + 27*8) + sum(4 + 16*n_v) + 27*8 per Gauss-Newton iteration."""
    return n_q * (16 + 27 * 8) + 4.0 * occupied + 16.0 * candidates + 27 * 8


class ClockSampler:
    """SM clock + throttle reasons sampled DURING the timed region: NVML from a thread every 5 ms (the timed region is tens
    of milliseconds, far too short for `nvidia-smi -lms`)."""
    BAD = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, device: int):
        import threading
        self.sm, self.reasons, self.max_mhz, self.stop_flag = [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            # NVML indexes physical devices: honour CUDA_VISIBLE_DEVICES when it lists plain ordinals
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = device
            if vis:
                try:
                    idx = int(vis.split(",")[device])
                except (ValueError, IndexError):
                    idx = device
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None
            return
        self.t = threading.Thread(target=self._run, daemon=True)
        self.t.start()

    def _run(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for bit, name in self.BAD.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        if self.nv is None:
            return out
        self.stop_flag = True
        self.t.join(timeout=2)
        if self.sm:
            out["sm_mhz"] = float(np.median(self.sm))
        out["reasons"] = sorted(self.reasons)
        out["samples"] = len(self.sm)
        return out


_OUT = None


def emit(text: str):
    out = _OUT if _OUT is not None else sys.stdout
    out.write(text + "\n")
    out.flush()


def host_cores() -> int:
    """Cores this process may run on — NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to its workers."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError, TypeError):
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s; MEASURED_PEAKS.json absent or without hbm_gbs)"


class ReferenceBuild:
    """The reference's OWN sage_icp::RegisterFrame (oracle/_ref/libsage_ref.so: the reference's hot-path sources compiled
    unmodified against the stand-in third-party headers of oracle/shim/, its tbb::parallel_reduce running on all host threads).
    It stops by its own rule (|log(est)| < 1e-4, at most 500 iterations), so a run is scaled to ITERS iterations with the
    iteration count of the oracle port on the same input — the two follow the same trajectory step for step
    (tests/test_reference_build.py).  None where the library did not travel."""

    def __init__(self, map_pts, threads):
        from oracle import ref_py
        self.map = None
        if not ref_py.available():
            return
        os.environ["SAGE_REF_THREADS"] = str(threads)
        self.map = ref_py.RefMap(VOXEL_SIZE_MAP, 1e9, BASIC, CRITICAL, BASIC_LABELS)
        self.map.add_points(map_pts)

    def seconds_per_scan(self, omap, sub, guess, n_full, threads):
        if self.map is None:
            return None, 0
        _, iters = omap.register_frame_core(sub, guess, MAX_DIST, KERNEL, SEM_TH, threads=threads)  # untimed: the iteration count
        t = time.perf_counter()
        self.map.register_frame_core(sub, guess, MAX_DIST, KERNEL, SEM_TH)
        dt = time.perf_counter() - t
        return dt * (ITERS / max(1, iters)) * (n_full / len(sub)), iters


def run_reference(args):
    """CPU arm, rank 0 only, all host threads: the reference algorithm's two CPU implementations available here — the oracle
    port (OpenMP over the two tbb::parallel_reduce sites) and, where it travelled, the reference's own code built against
    stand-in third-party headers (ReferenceBuild).  The line's value is the FASTER of the two (the more demanding baseline);
    cpu_baseline names it and carries both."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = host_cores()
    os.environ["OMP_NUM_THREADS"] = str(threads)  # before libgomp initialises (the library is loaded below); the calls also pass it
    from oracle import oracle_py as orc
    half = street_half_length(args.map_points)
    map_pts = make_map_points(args.map_points)
    omap = orc.OracleMap(VOXEL_SIZE_MAP, 1e9, BASIC, CRITICAL, BASIC_LABELS)
    omap.add_points(map_pts)
    rb = ReferenceBuild(map_pts, threads)
    # bounded sample: a fraction of the scan's queries so the whole run ends within minutes
    times, times_rb, iters_rb, frac = [], [], [], args.cpu_fraction
    for s in range(args.warmup + args.steps):
        scan, guess = make_queries(s, args.beams, args.az, half)
        sub = scan[:: max(1, int(round(1 / frac)))]
        t = time.perf_counter()
        omap.register_frame_core(sub, guess, MAX_DIST, KERNEL, SEM_TH, threads=threads, max_iters=ITERS, est_th=0.0)
        dt = time.perf_counter() - t
        t_rb, it = rb.seconds_per_scan(omap, sub, guess, len(scan), threads)
        if s >= args.warmup:
            times.append(dt * len(scan) / len(sub))  # scaled to the full scan
            if t_rb is not None:
                times_rb.append(t_rb), iters_rb.append(it)
    ms_port = 1e3 * float(np.mean(times))
    ms_rb = 1e3 * float(np.mean(times_rb)) if times_rb else None
    kind = "reference" if ms_rb is not None and ms_rb < ms_port else "port"
    ms = ms_rb if kind == "reference" else ms_port
    v = 1e3 / ms
    stride = max(1, int(round(1 / frac)))
    sample = (f"{args.steps} scans, " + ("every query" if stride == 1 else f"every {stride}-th query (time scaled to the full scan)")
              + f" of each {args.beams * args.az // 1000}k-pt scan x {ITERS} GN iters on {threads} host threads; value = the faster of: oracle "
              f"port (OpenMP) {1e3 / ms_port:.2f} scans/s"
              + (f", reference's own sources compiled against STAND-IN Eigen/Sophus/oneTBB/tsl headers (oracle/_ref; its parallel_reduce "
                 f"stand-in is a plain chunked thread pool, not TBB; runs of {int(np.mean(iters_rb))} iterations scaled to {ITERS}) "
                 f"{1e3 / ms_rb:.2f} scans/s" if ms_rb is not None else ", reference build not present"))
    emit(json.dumps({
        "impl": "reference", "metric": "RegisterFrame scans/sec (120 k-pt labeled scan)", "value": v, "unit": "scans/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, omap.num_points(), omap.num_voxels()),
        "cpu_baseline": {"value": v, "unit": "scans/s", "cores": threads, "kind": kind, "sample": sample,
                         "port_value": 1e3 / ms_port, "reference_build_value": (1e3 / ms_rb) if ms_rb is not None else None},
        "e2e": {"value": v, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def workload_config(args, map_points, map_voxels):
    rays = args.beams * args.az
    which = "configs[1]: 120k-pt" if rays == 120000 else f"configs[3]-style: {rays // 1000}k-pt dense"
    return {"workload": f"BASELINE {which} labelled scan as direct queries vs pre-built voxel map, 10 GN iterations "
                        "(kernel-level sage_icp::RegisterFrame)",
            "scan_rays": args.beams * args.az, "map_points": int(map_points), "map_voxels": int(map_voxels), "gn_iterations": ITERS,
            "max_correspondence_distance": MAX_DIST, "kernel": KERNEL, "sem_th": SEM_TH,
            "l2": "L2 flushed (256 MiB write) between timed steps; each step timed with its own CUDA-event pair",
            "parallelism": (f"query-shard x{args.gpus} (4 azimuth sectors per rank) + replicated map, all-reduce of the 17 normal-equation sums "
                            + ("fused into the search kernel over NVLink peer memory" if args.comm == "peer" else "by NCCL")) if args.gpus > 1 else "single GPU"}


def transform_by_guess(q: np.ndarray, g: np.ndarray) -> np.ndarray:
    """Positions the first Gauss-Newton iteration sees (yaw-only guess applied), for the occupancy statistics."""
    yaw = 2.0 * math.atan2(g[5], g[6])
    c, sn = math.cos(yaw), math.sin(yaw)
    qq = q.copy()
    qq[:, 0] = c * q[:, 0] - sn * q[:, 1] + g[0]
    qq[:, 1] = sn * q[:, 0] + c * q[:, 1] + g[1]
    qq[:, 2] = q[:, 2] + g[2]
    return qq


def pipeline_leg(sg, device: int, reps: int = 8):
    """BASELINE configs[0] at pipeline level: ONE 120 k-point labelled scan through sage_register_frame — the C-ABI call behind
    sageICP::RegisterFrame(frame): Preprocess + Voxelize + AdaptiveThreshold + ICP + map Update — from a PAGEABLE host buffer,
    against a map pre-built from 1 M surface samples, next to the oracle port of the same function single-threaded ("TBB off")
    and on all host cores; poses compared."""
    from oracle import oracle_py as orc
    from sage_icp_b200 import synthetic as syn
    cfg = sg.launch_config()
    world = make_map_points(1_000_000)
    local = world.copy()
    local[:, 2] -= syn.SENSOR_HEIGHT  # the pipeline's map frame is the first sensor frame: sensor at the origin
    truth = (0.3, 0.1, math.radians(1.0))
    scan = np.ascontiguousarray(syn.make_scan(4242, truth))  # plain numpy memory: pageable
    gp = sg.SagePipeline(cfg, device=device)
    ms, pose_g = [], None
    for r in range(reps + 2):
        gp.reinitialize()
        gp.map().add_points(local)
        n_map = gp.map().num_points()
        t = time.perf_counter()
        pose_g, _, _ = gp.register_frame(scan)
        dt = time.perf_counter() - t
        if r >= 2:
            ms.append(1e3 * dt)
    n_src, iters = len(gp.last_source()), gp.last_iterations()
    out = {"workload": "BASELINE configs[0]: one 120k-pt labelled scan through sage_register_frame (sageICP::RegisterFrame: Preprocess, Voxelize, "
                       "AdaptiveThreshold, ICP, map Update) from a pageable host buffer vs a map pre-built from 1 M samples, first frame "
                       "(identity initial guess, sigma = initial_threshold)",
           "scan_points": int(len(scan)), "map_points": int(n_map), "icp_queries_after_voxelize": int(n_src), "gn_iterations": int(iters),
           "gpu_ms_per_frame": float(np.median(ms)), "gpu_frames_per_s": 1e3 / float(np.median(ms)), "timing": f"wall clock around the call, median of {reps}"}
    cores = host_cores()
    for name, threads in (("cpu_single_thread", 1), ("cpu_all_cores", cores)):
        op = orc.OraclePipeline(cfg, threads=threads, evict_faithful=False)
        best, pose_c = None, None
        for r in range(2 if threads == 1 else 3):
            op.reset()
            op.map().add_points(local)
            t = time.perf_counter()
            pose_c, _, _ = op.register_frame(scan)
            dt = time.perf_counter() - t
            best = dt if best is None else min(best, dt)
        out[name] = {"ms_per_frame": 1e3 * best, "frames_per_s": 1.0 / best, "threads": threads, "kind": "port",
                     "iterations": op.last_iterations(), "queries": int(len(op.last_source()))}
        out["pose_delta_vs_cpu_m"] = float(np.linalg.norm(pose_g[:3] - pose_c[:3]))
        out["pose_delta_vs_cpu_rad"] = float(2.0 * min(np.linalg.norm(pose_g[3:] - pose_c[3:]), np.linalg.norm(pose_g[3:] + pose_c[3:])))
    out["speedup_vs_single_thread"] = out["cpu_single_thread"]["ms_per_frame"] / out["gpu_ms_per_frame"]
    out["speedup_vs_all_cores"] = out["cpu_all_cores"]["ms_per_frame"] / out["gpu_ms_per_frame"]
    return out


def hbm_regime_leg(sg, torch, device: int, map_points: int, base_pts: np.ndarray, base_half: float, peak: float, reps: int = 6):
    """BASELINE configs[4]'s DRAM-bound end: 120 k queries spread uniformly over a `map_points`-point map (no two queries share
    a bucket, nothing stays L2-resident), same kernel, 10 GN iterations.  The map is the bench map tiled along the street."""
    rng = np.random.default_rng(7)
    m = sg.SageMap(VOXEL_SIZE_MAP, 1e9, BASIC, CRITICAL, BASIC_LABELS, device=device)
    tiles = max(1, int(round(map_points / 5_000_000)))
    period = 2.0 * base_half + 40.0
    uq = []
    for t in range(tiles):
        pts = base_pts.copy()
        pts[:, 0] += t * period
        m.add_points(pts)
        uq.append(pts[rng.choice(len(pts), 120_000 // tiles + 1, replace=False)])
    q = np.concatenate(uq)[:120_000].copy()
    q[:, :3] += rng.normal(0, 0.1, (len(q), 3))  # near, not on, map points
    q = np.ascontiguousarray(q[rng.permutation(len(q))])
    ident = np.array([0, 0, 0, 0, 0, 0, 1.0])
    occ, cand = m.nn_stats(q)
    alg = algorithmic_bytes(len(q), occ, cand)
    work = m.search_work(q, MAX_DIST, SEM_TH, with_staged=True)
    d = torch.from_numpy(q).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        m.register_frame_device(d.data_ptr(), len(q), ident, MAX_DIST, KERNEL, SEM_TH, ITERS, 0.0)
    m.profile_enable(True)
    for _ in range(reps):
        flush.fill_(1)
        torch.cuda.synchronize()
        m.register_frame_device(d.data_ptr(), len(q), ident, MAX_DIST, KERNEL, SEM_TH, ITERS, 0.0)
    iters, launches, ms = m.profile_read_launches()
    m.profile_enable(False)
    us = 1e3 * ms / max(1, iters)
    # records the kernel pulls (staged by TMA in the tile search, else ranked straight from global memory) + table probes + query
    # in/out + winner record
    requested = 16.0 * ((work[4] if work[4] > 0 else work[0]) + work[1]) + 96.0 * len(q)
    out = {"workload": f"120k queries uniform over a {m.num_points()}-pt / {m.num_voxels()}-voxel map ({tiles} copies of the bench map along x), "
                       f"{ITERS} GN iterations, L2 flushed between registrations",
           "bound": "hbm", "algorithmic_bytes_per_iteration": alg, "us_per_iteration": us, "achieved": alg / us / 1e3, "peak": peak,
           "unit": "GB/s", "frac": alg / us / 1e3 / peak, "iterations_timed": int(iters), "launches_timed": int(launches),
           "requested_bytes_per_iteration": requested, "requested_gbs": requested / us / 1e3,
           "occupied_voxels_per_query": occ / len(q), "candidates_per_query": cand / len(q),
           "traffic": None, "traffic_source": None}
    tp = os.path.join(ROOT, "profiles", "traffic_hbm_regime.json")  # dram bytes per iteration from an ncu capture of this leg
    if os.path.exists(tp):
        tj = json.load(open(tp))
        out["traffic"], out["traffic_source"] = tj.get("dram_bytes_per_iteration"), tj.get("source")
    del m, d, flush
    torch.cuda.empty_cache()
    return out


def main():
    # stdout carries exactly one JSON line: anything a library prints there (NCCL's version banner under torchrun) is sent to
    # stderr instead, and the line itself goes to the saved descriptor
    global _OUT
    sys.stdout.flush()
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--map-points", type=int, default=5_000_000)
    ap.add_argument("--beams", type=int, default=64)
    ap.add_argument("--az", type=int, default=1875)
    ap.add_argument("--cpu-fraction", type=float, default=1.0, help="fraction of each scan the CPU arm times (1.0 = every query)")
    ap.add_argument("--no-pipeline", action="store_true", help="skip the configs[0] pipeline-level object")
    ap.add_argument("--no-hbm-regime", action="store_true", help="skip the 50 M-point uniform-query roofline object")
    ap.add_argument("--hbm-map-points", type=int, default=50_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--comm", default="peer", choices=["peer", "nccl"],
                    help="N > 1: all-reduce of the normal equations fused into the search kernel over NVLink peer memory, or NCCL")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import sage_icp_b200 as sg

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if sg.device_count() < 1:
        raise SystemExit("bench.py: no sm_100 device; sage_icp_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    # ---- untimed setup: replicated map, scans ------------------------------------------------------------
    half = street_half_length(args.map_points)
    map_pts = make_map_points(args.map_points)
    gmap = sg.SageMap(VOXEL_SIZE_MAP, 1e9, BASIC, CRITICAL, BASIC_LABELS, device=local)
    gmap.add_points(map_pts)
    n_map_points, n_map_voxels = gmap.num_points(), gmap.num_voxels()
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(sg.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        comm_used = args.comm
        if args.comm == "peer":
            # collective decision: if CUDA IPC / peer access fails on ANY rank, every rank falls back to NCCL
            ok = 1
            try:
                mine = torch.frombuffer(bytearray(gmap.comm_peer_handle()), dtype=torch.uint8).cuda()
            except sg.SageError as e:
                print(f"[rank {rank}] peer handle failed: {e}", file=sys.stderr)
                mine, ok = torch.zeros(64, dtype=torch.uint8, device="cuda"), 0
            allh = [torch.zeros(64, dtype=torch.uint8, device="cuda") for _ in range(world)]
            dist.all_gather(allh, mine)
            if ok:
                try:
                    gmap.comm_peer_attach(rank, world, b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
                except sg.SageError as e:
                    print(f"[rank {rank}] peer attach failed: {e}", file=sys.stderr)
                    ok = 0
            flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            if int(flag.item()) == 0:
                gmap.comm_destroy()
                comm_used = "nccl"
        if comm_used == "nccl":
            gmap.comm_init(rank, world, bytes(uid.cpu().numpy().tobytes()))
        args.comm = comm_used
        dist.barrier()

    total = args.warmup + args.steps
    scans, guesses, shards_dev, shards_pin = [], [], [], []
    for s in range(total):
        scan, guess = make_queries(s, args.beams, args.az, half)
        shard = shard_of(scan, rank, world, args.az)
        scans.append(scan); guesses.append(guess)
        shards_dev.append(torch.from_numpy(shard).cuda())
        shards_pin.append(torch.from_numpy(shard).pin_memory())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    stream = torch.cuda.ExternalStream(gmap.stream(), device=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(fn):
        """K steps after W warm-ups; every step bracketed by CUDA events on the library's stream; L2 flushed between."""
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for s in range(args.warmup):
            fn(s)
        barrier()
        l0 = sg.launch_count()
        wall0 = time.perf_counter()
        for k in range(args.steps):
            flush.fill_(k & 0xff)
            barrier()
            ev[k][0].record(stream)
            fn(args.warmup + k)
            ev[k][1].record(stream)
        barrier()
        wall = time.perf_counter() - wall0
        ms = [a.elapsed_time(b) for a, b in ev]
        tot = torch.tensor([sum(ms)], dtype=torch.float64, device="cuda")
        if dist is not None:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item()), sg.launch_count() - l0

    def step_resident(s):
        gmap.register_frame_device(shards_dev[s].data_ptr(), shards_dev[s].shape[0], guesses[s], MAX_DIST, KERNEL, SEM_TH, ITERS, 0.0)

    def step_e2e(s):
        gmap.register_frame(shards_pin[s].numpy(), guesses[s], MAX_DIST, KERNEL, SEM_TH, ITERS, 0.0)

    # ---- timed: resident ----------------------------------------------------------------------------------
    sampler = ClockSampler(local) if rank == 0 else None
    tot_ms, launches = timed_loop(step_resident)
    clocks = sampler.stop() if sampler else None
    # ---- timed: end to end through the host-buffer C-ABI call ---------------------------------------------
    e2e_ms, _ = timed_loop(step_e2e)
    # ---- roofline pass: per-kernel CUDA events inside the library (same steps) ----------------------------
    gmap.profile_enable(True)
    for s in range(args.warmup, total):
        flush.fill_(1)
        torch.cuda.synchronize()
        step_resident(s)
    n_iters, n_kernels, kernel_ms = gmap.profile_read_launches()
    gmap.profile_enable(False)
    alg, work = [], np.zeros(5)
    for s in range(args.warmup, total):
        # statistics at the positions the first iteration sees (guess applied), counted exactly on the device map
        qq = transform_by_guess(shards_dev[s].cpu().numpy(), guesses[s])
        occ, cand = gmap.nn_stats(qq)
        alg.append(algorithmic_bytes(len(qq), occ, cand))
        if s < args.warmup + 4:  # what the kernel really touches (first-iteration positions), a few steps are enough
            work += np.array(gmap.search_work(qq, MAX_DIST, SEM_TH, with_staged=True)) / len(qq)
            n_work = s - args.warmup + 1
    bytes_per_iter = float(np.mean(alg))
    iters_per_launch = n_iters / max(1, n_kernels)
    peak, peak_src = measured_peak_gbs()
    achieved = bytes_per_iter * n_iters / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes/launch of the same kernel from an ncu --set full capture
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic, traffic_src = tj.get("dram_bytes_per_launch"), tj.get("source")

    # ---- N > 1: the sharded registration must give the unsharded pose (rank 0 repeats it alone on its own GPU) ----------
    sharded_delta = None
    if world > 1:
        s = args.warmup
        pose_sh, _ = gmap.register_frame_device(shards_dev[s].data_ptr(), shards_dev[s].shape[0], guesses[s], MAX_DIST, KERNEL, SEM_TH, ITERS, 0.0)
        if rank == 0:
            solo = sg.SageMap(VOXEL_SIZE_MAP, 1e9, BASIC, CRITICAL, BASIC_LABELS, device=local)
            solo.add_points(map_pts)
            pose_1, _ = solo.register_frame(scans[s], guesses[s], MAX_DIST, KERNEL, SEM_TH, ITERS, 0.0)
            sharded_delta = float(np.linalg.norm(np.asarray(pose_sh) - np.asarray(pose_1)))
            del solo
            assert sharded_delta <= 1e-10, f"sharded pose differs from the unsharded one by {sharded_delta}"
        barrier()

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- CPU baseline on this box's host cores: the same scans, every query, all cores ---------------------
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        from oracle import oracle_py as orc
        threads = host_cores()
        omap = orc.OracleMap(VOXEL_SIZE_MAP, 1e9, BASIC, CRITICAL, BASIC_LABELS)
        omap.add_points(map_pts)
        stride = max(1, int(round(1 / args.cpu_fraction)))
        n_cpu = min(args.steps, 8)
        t_all, pose_c = [], None
        for k in range(n_cpu):
            sub = scans[args.warmup + k][::stride]
            t = time.perf_counter()
            pose_k, _ = omap.register_frame_core(sub, guesses[args.warmup + k], MAX_DIST, KERNEL, SEM_TH, threads=threads, max_iters=ITERS, est_th=0.0)
            t_all.append((time.perf_counter() - t) * len(scans[args.warmup + k]) / len(sub))
            pose_c = pose_k if k == 0 else pose_c
        t_all = float(np.mean(t_all))
        scan, guess = scans[args.warmup], guesses[args.warmup]
        sub = scan[::stride]
        sub1 = scan[::4]  # single thread ("TBB off"): a quarter of one scan, scaled
        t = time.perf_counter(); omap.register_frame_core(sub1, guess, MAX_DIST, KERNEL, SEM_TH, threads=1, max_iters=ITERS, est_th=0.0)
        t_one = (time.perf_counter() - t) * len(scan) / len(sub1)
        # parity check on the same scan through the GPU path
        pose_g, _ = gmap.register_frame(sub, guess, MAX_DIST, KERNEL, SEM_TH, ITERS, 0.0)
        try:  # the reference's own code, where its build travelled (extra information: never allowed to take the line down)
            t_rb, it_rb = ReferenceBuild(map_pts, threads).seconds_per_scan(omap, sub, guess, len(scan), threads)
        except Exception as e:  # noqa: BLE001
            print(f"reference build not timed: {e}", file=sys.stderr)
            t_rb, it_rb = None, 0
        kind = "reference" if t_rb is not None and t_rb < t_all else "port"
        sample = (f"{n_cpu} scans of the timed set, " + ("every query" if stride == 1 else f"every {stride}-th query, time scaled")
                  + f", x {ITERS} GN iters on the same {n_map_points}-pt map, {threads} host threads (mean); single_thread_value: a quarter of one scan, scaled")
        if t_rb is not None:
            sample += ("; value = the faster of the oracle port (OpenMP) and the reference's own sources compiled against STAND-IN "
                       f"Eigen/Sophus/oneTBB/tsl headers (oracle/_ref: one scan, a run of {it_rb} iterations scaled to {ITERS}; its parallel_reduce "
                       "stand-in is a plain chunked thread pool, not TBB)")
        cpu = {"value": 1.0 / (t_rb if kind == "reference" else t_all), "unit": "scans/s", "cores": threads, "kind": kind, "sample": sample,
               "port_value": 1.0 / t_all, "reference_build_value": (1.0 / t_rb) if t_rb is not None else None,
               "single_thread_value": 1.0 / t_one,
               "pose_delta_vs_gpu_m": float(np.linalg.norm(pose_g[:3] - pose_c[:3]))}

    # ---- pipeline level (configs[0]) and the DRAM-bound regime (configs[4]) -------------------------------------------
    pipeline = hbm = None
    if world == 1 and not args.no_pipeline:
        try:
            pipeline = pipeline_leg(sg, local)
        except Exception as e:  # noqa: BLE001 — an extra object must never take the line down
            pipeline = {"error": f"{type(e).__name__}: {e}"}
    if world == 1 and not args.no_hbm_regime and args.beams * args.az == 120000:
        try:
            del shards_dev, flush
            torch.cuda.empty_cache()
            hbm = hbm_regime_leg(sg, torch, local, args.hbm_map_points, map_pts, half, peak)
        except Exception as e:  # noqa: BLE001
            hbm = {"error": f"{type(e).__name__}: {e}"}

    n_scans = args.steps
    tile = work[4] > 0
    kernel_name = ("nn_tile_persistent_kernel" if tile and iters_per_launch > 1.5 else "nn_tile_kernel" if tile else "nn_search_kernel") + (
        " (correspondence search over TMA-staged buckets + normal equations + GN step; " if tile else
        " (correspondence search + normal equations + GN step; ") + (
        f"one cooperative launch runs the {iters_per_launch:.0f} iterations of a registration)" if iters_per_launch > 1.5 else "1 launch per iteration)")
    line = {
        "metric": "RegisterFrame scans/sec (120 k-pt labeled scan)", "value": n_scans / (tot_ms * 1e-3), "unit": "scans/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, n_map_points, n_map_voxels),
        "e2e": {"value": n_scans / (e2e_ms * 1e-3), "unit": "scans/s", "h2d_bytes_per_step": int(shards_pin[0].numel() * 8),
                "d2h_bytes_per_step": 7 * 8 + 4, "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": kernel_name,
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_iter * iters_per_launch, "launches_timed": int(n_kernels),
                     "avg_launch_us": 1e3 * kernel_ms / max(1, n_kernels),
                     "gn_iterations_per_launch": iters_per_launch, "algorithmic_bytes_per_iteration": bytes_per_iter,
                     "us_per_iteration": 1e3 * kernel_ms / max(1, n_iters),
                     "note": "achieved = algorithmic bytes of the reference's 27-voxel scan (SURVEY.md 8d) / measured kernel time; the kernel "
                             "prunes voxels that provably cannot hold the arg-min and shares staged buckets between the queries of a unit, so "
                             "it requests far fewer bytes than that (kernel_work_per_query); on this L2-resident working set the figure is an "
                             "algorithmic-throughput equivalence, not DRAM use — roofline_hbm_regime is the DRAM-bound measurement",
                     "kernel_work_per_query": {"records_ranked": work[0] / n_work, "table_probes": work[1] / n_work,
                                               "f64_reranked": work[2] / n_work,
                                               ("pooled_pairs" if tile else "deferred_to_warp_phase"): work[3] / n_work,
                                               "records_staged_by_tma": work[4] / n_work,
                                               "requested_bytes": (16 * (work[4] + work[1]) if tile else 16 * (work[0] + work[1])) / n_work + 32 + 32 + 32}},
        "cpu_baseline": cpu,
        "pipeline": pipeline,
        "roofline_hbm_regime": hbm,
    }
    if sharded_delta is not None:
        line["sharded_pose_delta_m"] = sharded_delta
    emit(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
