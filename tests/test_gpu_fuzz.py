"""Randomised parity: random map geometry / capacities / label sets / thresholds / scene scale, GPU vs oracle.  Every case
checks the three order-sensitive stages bit for bit (map contents after sequential AddPoint semantics, per-query
correspondences, down-sampled clouds in robin_map order) and one registration within the pose tolerance."""
import numpy as np
import pytest

from conftest import POSE_TOL_M, POSE_TOL_RAD, assert_maps_equal, pose_delta

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("search_mode")]

LABEL_POOL = [0, 10, 11, 40, 44, 48, 49, 50, 51, 70, 71, 72, 80, 81, 99, 252]


@pytest.mark.parametrize("seed", range(24))
def test_random_configuration_parity(orc, seed):
    import sage_icp_b200 as sg
    from sage_icp_b200.config import SageConfig
    rng = np.random.default_rng(1000 + seed)
    vs_map = float(rng.choice([0.3, 0.5, 0.8, 1.0, 1.7]))
    basic, critical = int(rng.integers(1, 25)), int(rng.integers(0, 25))
    basic_labels = [int(l) for l in rng.choice(LABEL_POOL[1:], size=int(rng.integers(1, 6)), replace=False)]
    sem_th = float(rng.choice([0.05, 0.2, 0.4, 0.8, 1.0, 1.5]))
    scale = float(rng.choice([3.0, 10.0, 40.0]))
    offset = rng.uniform(-1, 1, 3) * float(rng.choice([0.0, 50.0, 5000.0]))
    n_map, n_q = int(rng.integers(2000, 40000)), int(rng.integers(50, 6000))

    # clustered scene: a few planes and blobs so that voxels fill up and neighbourhoods vary
    centres = rng.uniform(-scale, scale, (12, 3))
    pts = centres[rng.integers(0, 12, n_map)] + rng.normal(0, scale * 0.08, (n_map, 3)) * rng.choice([[1, 1, 0.02], [1, 0.02, 1], [1, 1, 1]], n_map)
    pts = np.c_[pts + offset, rng.choice(LABEL_POOL, n_map).astype(float)]
    pts[:, :3] = pts[:, :3].astype(np.float32)
    g = sg.SageMap(vs_map, 1e9, basic, critical, basic_labels)
    o = orc.OracleMap(vs_map, 1e9, basic, critical, basic_labels, evict_faithful=False)
    for chunk in np.array_split(pts, int(rng.integers(1, 4))):
        g.add_points(chunk)
        o.add_points(chunk)
    assert_maps_equal(g.dump(), o.dump())

    q = pts[rng.integers(0, n_map, n_q)].copy()
    q[:, :3] += rng.normal(0, vs_map * rng.choice([0.05, 0.5, 2.0]), (n_q, 3))
    q[:, 3] = rng.choice(LABEL_POOL, n_q)
    max_dist = float(vs_map * rng.choice([0.5, 1.5, 4.0]))
    tgt, matched = g.get_correspondences(q, max_dist, sem_th)
    _, tgt_o, qidx = o.get_correspondences(q, max_dist, sem_th)
    m_o = np.zeros(n_q, bool)
    m_o[qidx] = True
    assert np.array_equal(matched, m_o) and np.array_equal(tgt[matched], tgt_o)

    # registration from a small perturbation (skip degenerate cases: too few pairs make the 6x6 system singular)
    if matched.sum() > 200:
        guess = orc.se3_exp(rng.normal(0, 1, 6) * [0.05, 0.05, 0.05, 0.002, 0.002, 0.002])
        kern = float(rng.choice([0.1, 0.33, 1.0]))
        pose_o, it_o = o.register_frame_core(q, guess, max_dist, kern, sem_th, max_iters=30)
        pose_g, it_g = g.register_frame(q, guess, max_dist, kern, sem_th, max_iters=30)
        dt, da = pose_delta(pose_g, pose_o)
        assert it_g == it_o and dt <= POSE_TOL_M and da <= POSE_TOL_RAD, (it_g, it_o, dt, da)

    # front end with a random grouping of the labels
    labels = [int(l) for l in rng.permutation(LABEL_POOL[:-1])]
    cuts = sorted(rng.choice(np.arange(1, len(labels)), size=int(rng.integers(1, 5)), replace=False))
    groups = [labels[a:b] for a, b in zip([0] + list(cuts), list(cuts) + [len(labels)])]
    cfg = SageConfig(voxel_labels=groups, voxel_size=[float(rng.choice([0.3, 0.6, 1.0, 2.0])) for _ in groups], voxel_size_map=vs_map,
                     max_range=float(scale * 3), min_range=float(scale * 0.05), label_max_range=float(scale), basic_points_per_voxel=basic,
                     critical_points_per_voxel=critical, basic_parts_labels=basic_labels, sem_th=sem_th, dynamic_vehicle_voxid=0)
    p = sg.SagePipeline(cfg)
    local = q.copy()
    local[:, :3] -= offset
    local[:, :3] = local[:, :3].astype(np.float32)
    cropped = orc.preprocess(local, cfg.max_range, cfg.min_range, cfg.label_max_range)
    assert np.array_equal(p.preprocess(local), cropped)
    for s in (0.5, 1.5):
        assert np.array_equal(p.voxel_downsample(cropped, s), orc.voxel_downsample(cfg, cropped, s)), s
