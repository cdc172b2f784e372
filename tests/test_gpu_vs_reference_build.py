"""GPU parity against the REFERENCE'S OWN CODE (oracle/_ref/libsage_ref.so: the reference's unmodified hot-path sources compiled
against the stand-in third-party headers of oracle/shim/, see tests/test_reference_build.py) — the CUDA path and the reference's
C++ on the same seeded inputs, without the oracle in between.  The library is built in the development container (where
/root/reference is) and travels to the GPU box with the snapshot; without it these tests skip."""
import numpy as np
import pytest

from conftest import POSE_TOL_M, POSE_TOL_RAD, pose_delta

pytestmark = pytest.mark.gpu

BASIC_LABELS = [40, 44, 48, 49, 50, 70, 72]


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_py
    if not ref_py.available():
        pytest.skip("oracle/_ref/libsage_ref.so did not travel and /root/reference is not here")
    ref_py.lib()
    return ref_py


def test_correspondences_equal_the_reference(ref):
    """VoxelHashMap::GetCorrespondences of the reference itself vs nn_search_kernel: the same pairs, bit for bit."""
    import sage_icp_b200 as sg
    from sage_icp_b200 import synthetic as syn
    pts = syn.sample_street_map(150_000, 7, -40.0, 40.0)
    r = ref.RefMap(0.8, 100.0, 20, 20, BASIC_LABELS)
    r.add_points(pts)
    g = sg.SageMap(0.8, 100.0, 20, 20, BASIC_LABELS)
    g.add_points(pts)
    q = syn.make_scan(11, (0.0, 0.0, 0.0), n_beams=32, n_az=400).copy()
    q[:, 2] += 1.73
    q[:, 0] += 0.3
    rs, rt = r.get_correspondences(q, 2.0, 0.4)
    tg, mg = g.get_correspondences(q, 2.0, 0.4)
    mg = mg.astype(bool)
    assert mg.sum() == len(rs) > 1000
    assert np.array_equal(q[mg], rs) and np.array_equal(tg[mg], rt)


def test_register_frame_drive_equals_the_reference(ref):
    """sageICP::RegisterFrame of the reference itself vs the CUDA pipeline (reference map semantics switched on, so the local map
    can be compared in order): poses within the north-star tolerance (observed: rounding level), identical query clouds."""
    import sage_icp_b200 as sg
    from sage_icp_b200 import synthetic as syn
    from sage_icp_b200.config import launch_config
    cfg = launch_config()
    gp, rp = sg.SagePipeline(cfg), ref.RefPipeline(cfg)
    gp.map().set_eviction(True)
    n = 20
    traj = syn.trajectory(n)
    worst = 0.0
    for i in range(n):
        scan = syn.make_scan(100 + i, tuple(traj[i]), n_beams=32, n_az=900)
        pg, _, _ = gp.register_frame(scan)
        pr = rp.register_frame(scan)
        dt, da = pose_delta(pg, pr)
        worst = max(worst, dt)
        assert dt <= POSE_TOL_M and da <= POSE_TOL_RAD, (i, dt, da)
        assert np.array_equal(gp.last_source(), rp.last_source()), i
    assert worst < 1e-6
    a, b = gp.local_map(), rp.local_map()
    assert a.shape == b.shape and np.array_equal(a[:, 3], b[:, 3]) and np.allclose(a, b, atol=1e-6, rtol=0)


@pytest.mark.parametrize("overrides", [dict(), dict(voxel_size_map=1.0, sem_th=0.8), dict(sem_th=0.2)], ids=["odometry", "odometry_360", "odometry_raw"])
def test_dynamic_vehicle_filter_drive_equals_the_reference(ref, overrides):
    """dynamic_vehicle_filter = true (default in three of the four launch files): the reference's own Preprocess — cluster by
    cluster emission over the stand-in PCL — vs the CUDA front end, frame by frame: identical filtered clouds in order, identical
    query clouds, poses within tolerance (core/Preprocessing.cpp:95-172)."""
    import sage_icp_b200 as sg
    from sage_icp_b200 import synthetic as syn
    from sage_icp_b200.config import launch_config
    cfg = launch_config(dynamic_vehicle_filter=True, **overrides)
    gp, rp = sg.SagePipeline(cfg), ref.RefPipeline(cfg)
    n = 12
    traj = syn.trajectory(n)
    for i in range(n):
        scan = syn.make_scan(700 + i, tuple(traj[i]), n_beams=48, n_az=1000)
        if i < 3:
            assert np.array_equal(gp.preprocess(scan), ref.preprocess(cfg, scan)), i
        pg, _, _ = gp.register_frame(scan)
        pr = rp.register_frame(scan)
        dt, da = pose_delta(pg, pr)
        assert dt <= POSE_TOL_M and da <= POSE_TOL_RAD, (i, dt, da)
        assert np.array_equal(gp.last_source(), rp.last_source()), i
