#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the ORACLE (oracle/, the CPU restatement of the reference).  The reference's own sources,
built against stand-in third-party headers (oracle/_ref), reproduce these files: tests/test_reference_build.py.  The fixtures pin (a) the oracle against
silent regressions and (b) the CUDA path on the GPU box, where /root/reference and this generator's inputs need not exist.
Run:  python tools/make_golden.py        (deterministic; commit the result)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_py as orc  # noqa: E402
from sage_icp_b200 import synthetic as syn  # noqa: E402
from sage_icp_b200.config import launch_config  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
BASIC_LABELS = [40, 44, 48, 49, 50, 70, 72]


def pack_scan(scan):
    """Scans are f32 xyz + small-integer labels widened to f64: store them that way, losslessly."""
    assert np.array_equal(scan[:, :3], scan[:, :3].astype(np.float32).astype(np.float64))
    return scan[:, :3].astype(np.float32), scan[:, 3].astype(np.int16)


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261017)
    cfg = launch_config()

    # 1. SE(3): exp / log / compose
    xi = rng.normal(size=(32, 6)) * np.array([2, 2, 2, 0.6, 0.6, 0.6])
    xi[0] = 0
    xi[1, 3:] = 1e-12
    poses = np.array([orc.se3_exp(x) for x in xi])
    np.savez_compressed(os.path.join(OUT, "se3.npz"), xi=xi, exp=poses, log=np.array([orc.se3_log(p) for p in poses]),
                        mul=np.array([orc.se3_mul(poses[i], poses[(i + 1) % 32]) for i in range(32)]),
                        inv=np.array([orc.se3_inverse(p) for p in poses]))

    # 2. Preprocess + VoxelDownsample (reference output order) on one small labelled scan
    scan = syn.make_scan(7, (0.0, 0.0, 0.0), n_beams=16, n_az=450)
    xyz, lab = pack_scan(scan)
    cropped = orc.preprocess(scan, cfg.max_range, cfg.min_range, cfg.label_max_range)
    ds = orc.voxel_downsample(cfg, cropped, 0.5)
    src = orc.voxel_downsample(cfg, ds, 1.5)
    np.savez_compressed(os.path.join(OUT, "frontend.npz"), xyz=xyz, label=lab, cropped=cropped, downsample=ds, source=src)

    # 3. map build (AddPoint rule) + GetCorrespondences + normal equations
    pts = syn.sample_street_map(12_000, 11, -12.0, 12.0)
    pts = np.c_[pts[:, :3].astype(np.float32).astype(np.float64), pts[:, 3]]
    m = orc.OracleMap(0.8, 100.0, 20, 20, BASIC_LABELS, evict_faithful=False)
    m.add_points(pts)
    keys, counts, vox = m.dump()
    q = syn.make_scan(9, (0.0, 0.0, 0.0), n_beams=16, n_az=200)
    r = np.linalg.norm(q[:, :3], axis=1)
    q = q[(r > 3) & (r < 14)]
    q[:, :3] = (q[:, :3] + np.array([0.25, -0.1, syn.SENSOR_HEIGHT])).astype(np.float32)
    _, tgt, qidx = m.get_correspondences(q, 1.5, 0.4)
    s_o, t_o, _ = m.get_correspondences(q, 1.5, 0.4)
    JTJ, JTr, x, est = orc.align_clouds(s_o, t_o, 0.5)
    guess = syn.pose7_from_xyyaw((0.1, 0.05, 0.004))
    scan_local = q.copy()
    pose, iters = m.register_frame_core(scan_local, guess, 3.0, 1.0 / 3.0, 0.4)
    order = np.lexsort(keys.T[::-1])
    np.savez_compressed(os.path.join(OUT, "core.npz"), map_xyz=pts[:, :3].astype(np.float32), map_label=pts[:, 3].astype(np.int16),
                        keys=keys[order], counts=counts[order], voxels=vox[order].astype(np.float32), queries=q, matched_idx=qidx, targets=tgt,
                        JTJ=JTJ, JTr=JTr, x=x, guess=guess, pose=pose, iters=iters)

    # 4. a short RegisterFrame sequence (pipeline level)
    traj = syn.trajectory(5)
    p = orc.OraclePipeline(cfg, evict_faithful=False)
    xyzs, labs, out_pose, sig, its, nsrc, nds = [], [], [], [], [], [], []
    for i in range(5):
        scan = syn.make_scan(300 + i, tuple(traj[i]), n_beams=16, n_az=450)
        a, b = pack_scan(scan)
        xyzs.append(a); labs.append(b)
        pose, _, _ = p.register_frame(scan)
        out_pose.append(pose); sig.append(p.last_sigma()); its.append(p.last_iterations())
        nsrc.append(len(p.last_source())); nds.append(len(p.last_frame_downsample()))
    np.savez_compressed(os.path.join(OUT, "sequence.npz"), xyz=np.array(xyzs), label=np.array(labs), poses=np.array(out_pose),
                        sigma=np.array(sig), iterations=np.array(its), n_source=np.array(nsrc), n_downsample=np.array(nds),
                        map_voxels=p.map().num_voxels(), map_points=p.map().num_points())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
