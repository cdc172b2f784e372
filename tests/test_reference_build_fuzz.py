"""Randomised oracle-vs-reference-build parity: the SAME 24 random configurations as tests/test_gpu_fuzz.py (map geometry,
capacities, label sets, thresholds, scene scale and offset), here between the oracle and the reference's own code
(oracle/_ref, see tests/test_reference_build.py).  Together the two files close the chain CUDA path == oracle == reference code on
identical inputs.  Order-sensitive stages bit for bit; the registration to rounding (it runs to the reference's own stopping
rule, which the oracle's defaults restate)."""
import numpy as np
import pytest

from conftest import pose_delta

LABEL_POOL = [0, 10, 11, 40, 44, 48, 49, 50, 51, 70, 71, 72, 80, 81, 99, 252]


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_py
    if not ref_py.available():
        pytest.skip("neither oracle/_ref/libsage_ref.so nor /root/reference is present")
    ref_py.lib()
    return ref_py


@pytest.mark.parametrize("seed", range(24))
def test_random_configuration_parity(ref, orc, seed):
    from sage_icp_b200.config import SageConfig
    rng = np.random.default_rng(1000 + seed)  # the same stream of draws as tests/test_gpu_fuzz.py
    vs_map = float(rng.choice([0.3, 0.5, 0.8, 1.0, 1.7]))
    basic, critical = int(rng.integers(1, 25)), int(rng.integers(0, 25))
    basic_labels = [int(l) for l in rng.choice(LABEL_POOL[1:], size=int(rng.integers(1, 6)), replace=False)]
    sem_th = float(rng.choice([0.05, 0.2, 0.4, 0.8, 1.0, 1.5]))
    scale = float(rng.choice([3.0, 10.0, 40.0]))
    offset = rng.uniform(-1, 1, 3) * float(rng.choice([0.0, 50.0, 5000.0]))
    n_map, n_q = int(rng.integers(2000, 40000)), int(rng.integers(50, 6000))

    centres = rng.uniform(-scale, scale, (12, 3))
    pts = centres[rng.integers(0, 12, n_map)] + rng.normal(0, scale * 0.08, (n_map, 3)) * rng.choice([[1, 1, 0.02], [1, 0.02, 1], [1, 1, 1]], n_map)
    pts = np.c_[pts + offset, rng.choice(LABEL_POOL, n_map).astype(float)]
    pts[:, :3] = pts[:, :3].astype(np.float32)
    r = ref.RefMap(vs_map, 1e9, basic, critical, basic_labels)
    o = orc.OracleMap(vs_map, 1e9, basic, critical, basic_labels, evict_faithful=True)
    for chunk in np.array_split(pts, int(rng.integers(1, 4))):
        r.add_points(chunk)
        o.add_points(chunk)
    (rk, rc, rp), (ok, oc, op) = r.dump(), o.dump()
    assert np.array_equal(rk, ok) and np.array_equal(rc, oc) and np.array_equal(rp, op)  # voxels, map order, stored order

    q = pts[rng.integers(0, n_map, n_q)].copy()
    q[:, :3] += rng.normal(0, vs_map * rng.choice([0.05, 0.5, 2.0]), (n_q, 3))
    q[:, 3] = rng.choice(LABEL_POOL, n_q)
    max_dist = float(vs_map * rng.choice([0.5, 1.5, 4.0]))
    rs, rt = r.get_correspondences(q, max_dist, sem_th)
    os_, ot, qidx = o.get_correspondences(q, max_dist, sem_th)
    assert np.array_equal(rs, os_) and np.array_equal(rt, ot) and np.array_equal(rs, q[qidx])

    if len(qidx) > 200:
        guess = orc.se3_exp(rng.normal(0, 1, 6) * [0.05, 0.05, 0.05, 0.002, 0.002, 0.002])
        kern = float(rng.choice([0.1, 0.33, 1.0]))
        pose_o, it_o = o.register_frame_core(q, guess, max_dist, kern, sem_th)  # defaults = the reference's 500 / 1e-4
        pose_r = r.register_frame_core(q, guess, max_dist, kern, sem_th)
        dt, da = pose_delta(pose_r, pose_o)
        assert dt < 1e-7 and da < 1e-8, (it_o, dt, da)

    labels = [int(l) for l in rng.permutation(LABEL_POOL[:-1])]
    cuts = sorted(rng.choice(np.arange(1, len(labels)), size=int(rng.integers(1, 5)), replace=False))
    groups = [labels[a:b] for a, b in zip([0] + list(cuts), list(cuts) + [len(labels)])]
    cfg = SageConfig(voxel_labels=groups, voxel_size=[float(rng.choice([0.3, 0.6, 1.0, 2.0])) for _ in groups], voxel_size_map=vs_map,
                     max_range=float(scale * 3), min_range=float(scale * 0.05), label_max_range=float(scale), basic_points_per_voxel=basic,
                     critical_points_per_voxel=critical, basic_parts_labels=basic_labels, sem_th=sem_th, dynamic_vehicle_voxid=0)
    local = q.copy()
    local[:, :3] -= offset
    local[:, :3] = local[:, :3].astype(np.float32)
    cropped = orc.preprocess(local, cfg.max_range, cfg.min_range, cfg.label_max_range)
    assert np.array_equal(ref.preprocess(cfg, local), cropped)
    for s in (0.5, 1.5):
        assert np.array_equal(ref.voxel_downsample(cfg, cropped, s), orc.voxel_downsample(cfg, cropped, s)), s
