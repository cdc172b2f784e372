"""CPU-only: the streaming workload of tools/stream_bench.py (BASELINE configs[2]: full sageICP::RegisterFrame over a synthetic
KITTI-shaped drive, 64 x 1875-ray scans, map built incrementally) timed for the two CPU implementations of the reference algorithm
available here — the oracle port (OpenMP) and the reference's own code (oracle/_ref: the reference's sources against stand-in
third-party headers, SAGE_REF_THREADS threads in its parallel_reduce) — with their poses compared.  Test infrastructure.

    python tools/cpu_stream_bench.py [--frames 60] [--threads N]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=60)
    ap.add_argument("--threads", type=int, default=0, help="0 = all host threads")
    a = ap.parse_args()
    from oracle import oracle_py as orc, ref_py as ref
    from sage_icp_b200 import synthetic as syn
    from sage_icp_b200.config import launch_config
    threads = a.threads or orc.max_threads()
    os.environ["SAGE_REF_THREADS"] = str(threads)
    cfg = launch_config()
    op, rp = orc.OraclePipeline(cfg, threads=threads, evict_faithful=True), ref.RefPipeline(cfg)
    traj = syn.trajectory(a.frames)
    t_o, t_r, worst = [], [], 0.0
    for i in range(a.frames):
        scan = syn.make_scan(i, tuple(traj[i]), n_beams=64, n_az=1875)
        t = time.perf_counter(); po, _, _ = op.register_frame(scan); t_o.append(time.perf_counter() - t)
        t = time.perf_counter(); pr = rp.register_frame(scan); t_r.append(time.perf_counter() - t)
        worst = max(worst, float(np.linalg.norm(po[:3] - pr[:3])))
    w = min(5, a.frames // 2)
    print(json.dumps({"workload": "BASELINE configs[2] streaming RegisterFrame, CPU only", "frames": a.frames, "threads": threads,
                      "host": os.popen("lscpu | grep 'Model name' | sed 's/.*: *//'").read().strip(),
                      "oracle_port_frames_per_s": 1.0 / float(np.mean(t_o[w:])), "reference_build_frames_per_s": 1.0 / float(np.mean(t_r[w:])),
                      "max_pose_delta_m": worst, "mean_queries": float(len(op.last_source()))}))


if __name__ == "__main__":
    main()
