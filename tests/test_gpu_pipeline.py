"""GPU parity tests of the pipeline level (Preprocess, VoxelDownsample, Voxelize, RegisterFrame) vs the oracle."""
import numpy as np
import pytest

from conftest import POSE_TOL_M, POSE_TOL_RAD, assert_maps_equal, pose_delta

pytestmark = pytest.mark.gpu


def _scan(seed, pose=(0.0, 0.0, 0.0), beams=32, az=600):
    from sage_icp_b200 import synthetic as syn
    return syn.make_scan(seed, pose, n_beams=beams, n_az=az)


def test_preprocess_bit_exact(orc, cfg):
    import sage_icp_b200 as sg
    p = sg.SagePipeline(cfg)
    scan = _scan(1)
    out = p.preprocess(scan)
    ref = orc.preprocess(scan, cfg.max_range, cfg.min_range, cfg.label_max_range)
    assert out.shape == ref.shape and np.array_equal(out, ref)
    assert (ref[:, 3] == 0).sum() > (scan[:, 3] == 0).sum()  # far labels were zeroed
    assert len(p.preprocess(np.zeros((0, 4)))) == 0


@pytest.mark.parametrize("scale", [0.5, 1.5, 1.0])
@pytest.mark.parametrize("beams,az", [(32, 600), (64, 1875), (4, 30)])
def test_voxel_downsample_bit_exact_with_reference_order(orc, cfg, scale, beams, az):
    """Same survivors AND same output order (robin_map iteration order per group, groups concatenated)."""
    import sage_icp_b200 as sg
    p = sg.SagePipeline(cfg)
    scan = orc.preprocess(_scan(2, beams=beams, az=az), cfg.max_range, cfg.min_range, cfg.label_max_range)
    out = p.voxel_downsample(scan, scale)
    ref = orc.voxel_downsample(cfg, scan, scale)
    assert out.shape == ref.shape
    assert np.array_equal(out, ref)
    # points whose label is in no group were dropped
    assert not np.isin(out[:, 3], [30, 252]).any()


def test_voxelize_matches_oracle(orc, cfg):
    import sage_icp_b200 as sg
    p = sg.SagePipeline(cfg)
    op = orc.OraclePipeline(cfg)
    scan = orc.preprocess(_scan(3), cfg.max_range, cfg.min_range, cfg.label_max_range)
    s, d = p.voxelize(scan)
    so, do = op.voxelize(scan)
    assert np.array_equal(d, do) and np.array_equal(s, so)


@pytest.mark.parametrize("variant", ["odometry", "360", "raw"])
def test_register_frame_sequence_parity(orc, variant):
    """Full sageICP::RegisterFrame over a short drive: per-frame pose parity, identical query clouds, identical
    sigma / iteration counts, identical map at the end (oracle in clean-eviction mode)."""
    import sage_icp_b200 as sg
    from sage_icp_b200 import synthetic as syn
    from sage_icp_b200.config import launch_config
    over = {"odometry": {}, "360": dict(voxel_size_map=1.0, sem_th=0.8), "raw": dict(sem_th=0.2, local_map_range=40.0)}[variant]
    cfg = launch_config(**over)
    gp = sg.SagePipeline(cfg)
    op = orc.OraclePipeline(cfg, evict_faithful=False)
    n = 25
    traj = syn.trajectory(n)
    for i in range(n):
        scan = syn.make_scan(100 + i, tuple(traj[i]), n_beams=32, n_az=900)
        pg, ti, ta = gp.register_frame(scan)
        po, _, _ = op.register_frame(scan)
        dt, da = pose_delta(pg, po)
        assert dt <= POSE_TOL_M and da <= POSE_TOL_RAD, (i, dt, da)
        assert gp.last_iterations() == op.last_iterations(), i
        assert gp.last_sigma() == pytest.approx(op.last_sigma(), rel=1e-9, abs=1e-12)
        assert np.array_equal(gp.last_source(), op.last_source()), i
        assert np.array_equal(gp.last_frame_downsample(), op.last_frame_downsample()), i
        assert 0 <= ti <= ta
    assert np.abs(gp.poses() - op.poses()).max() < 1e-6
    gm, om = gp.map().dump(), op.map().dump()
    # map points are pose * point with poses that agree to ~1e-12, so compare with a tolerance
    from conftest import map_as_dict
    dg, do = map_as_dict(*gm), map_as_dict(*om)
    assert set(dg) == set(do)
    for k in dg:
        assert dg[k].shape == do[k].shape
        assert np.allclose(dg[k], do[k], atol=1e-7, rtol=0)
    # LocalMap(): same point set
    a, b = gp.local_map(), op.local_map()
    assert a.shape == b.shape


def test_reinitialize(orc, cfg):
    import sage_icp_b200 as sg
    gp = sg.SagePipeline(cfg)
    for i in range(3):
        gp.register_frame(_scan(i, (0.2 * i, 0, 0)))
    assert len(gp.poses()) == 3 and gp.map().num_voxels() > 0
    gp.reinitialize()
    assert len(gp.poses()) == 0 and gp.map().empty()
    pose, _, _ = gp.register_frame(_scan(0))
    assert np.allclose(pose, [0, 0, 0, 0, 0, 0, 1])


def test_set_devices(orc, cfg):
    """sage_set_devices: a no-op for the current device; refused for unusable or repeated GPUs and once the pipeline holds state;
    moving a fresh pipeline to another GPU (when the box has one) gives the same poses."""
    import sage_icp_b200 as sg
    gp = sg.SagePipeline(cfg)
    gp.set_devices([0])
    assert gp.num_devices() == 1
    with pytest.raises(sg.SageError, match="twice"):
        gp.set_devices([0, 0])
    with pytest.raises(sg.SageError):
        gp.set_devices([sg.device_count()])  # no such device; the handle keeps working on the old one
    with pytest.raises(sg.SageError):
        gp.set_devices([0, sg.device_count()])
    ref = [gp.register_frame(_scan(i, (0.2 * i, 0, 0)))[0] for i in range(3)]
    with pytest.raises(sg.SageError, match="fresh"):
        gp.set_devices([1])
    if sg.device_count() >= 2:
        other = sg.SagePipeline(cfg)
        other.set_devices([1])
        for i in range(3):
            assert np.array_equal(other.register_frame(_scan(i, (0.2 * i, 0, 0)))[0], ref[i])


def test_one_handle_on_two_gpus(orc, cfg):
    """sage_set_devices(h, {0, 1}, 2): ONE process, the ICP queries of every frame sharded over both GPUs, the sums all-reduced
    inside the search kernel over peer memory, replicated map updates.  Poses equal the single-GPU ones to rounding (another
    summation order), the oracle's within the tolerance; the query clouds are identical; reinitialize() clears every replica."""
    import sage_icp_b200 as sg
    from sage_icp_b200 import synthetic as syn
    if sg.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    one, two, op = sg.SagePipeline(cfg), sg.SagePipeline(cfg), orc.OraclePipeline(cfg, threads=orc.max_threads(), evict_faithful=False)
    two.set_devices([0, 1])
    assert two.num_devices() == 2
    n = 25
    traj = syn.trajectory(n)
    for rnd in range(2):
        for i in range(n):
            scan = syn.make_scan(800 + i, tuple(traj[i]))
            p1, _, _ = one.register_frame(scan)
            p2, _, _ = two.register_frame(scan)
            po, _, _ = op.register_frame(scan)
            assert float(np.abs(p1 - p2).max()) <= 1e-10, (rnd, i, p1, p2)
            dt, da = pose_delta(p2, po)
            assert dt <= POSE_TOL_M and da <= POSE_TOL_RAD, (rnd, i, dt, da)
            assert two.last_iterations() == one.last_iterations() == op.last_iterations(), (rnd, i)
            assert np.array_equal(two.last_source(), op.last_source()), (rnd, i)
        assert two.map().num_voxels() == one.map().num_voxels() == op.map().num_voxels()
        one.reinitialize(), two.reinitialize(), op.reset()
        assert two.map().num_voxels() == 0 and len(two.poses()) == 0
    with pytest.raises(sg.SageError, match="fresh"):
        two.register_frame(syn.make_scan(1, (0.0, 0.0, 0.0))), two.set_devices([0])


def test_bad_config_fails_loudly(cfg):
    import sage_icp_b200 as sg
    from sage_icp_b200.config import SageConfig, launch_config
    with pytest.raises(sg.SageError):
        sg.SagePipeline(SageConfig())  # empty voxel_labels: UB in the reference, error here
    with pytest.raises(sg.SageError):
        sg.SagePipeline(launch_config(dynamic_vehicle_voxid=9))  # voxel_labels[9]: out of range (UB in the reference)


@pytest.mark.parametrize("dy_th,seed", [(0.5, 3), (0.05, 4), (3.0, 5)])
def test_dynamic_vehicle_filter_preprocess(orc, dy_th, seed):
    """Preprocess with dynamic_vehicle_filter = true (core/Preprocessing.cpp:95-172): same kept points in the same order as the
    oracle's restatement of the PCL clustering + landmark test in the reference's emission order — plain inliers in input order,
    then the kept vehicle clusters by descending size (ties: smallest member first), members ascending — which the reference
    build reproduces point for point (tests/test_reference_build.py)."""
    import sage_icp_b200 as sg
    from sage_icp_b200.config import launch_config
    cfg = launch_config(dynamic_vehicle_filter=True, dynamic_vehicle_filter_th=dy_th)
    p = sg.SagePipeline(cfg)
    scan = _scan(seed, beams=48, az=1000)
    out = p.preprocess(scan)
    ref = orc.preprocess_dynamic(cfg, scan, cluster_order=True)
    assert out.shape == ref.shape and np.array_equal(out, ref)
    by_input = orc.preprocess_dynamic(cfg, scan, cluster_order=False)
    if dy_th == 0.5:
        assert not np.array_equal(ref, by_input)  # the cluster order really differs from input order on this scan
    plain = orc.preprocess(scan, cfg.max_range, cfg.min_range, cfg.label_max_range)
    veh = [10, 11, 13, 15, 16, 18, 20]
    n_in, n_out = np.isin(plain[:, 3], veh).sum(), np.isin(out[:, 3], veh).sum()
    assert n_in > 500 and n_out <= n_in
    if dy_th == 0.5:
        assert 0 < n_out < n_in  # some clusters stay (next to sidewalk / parking points), some go
    # non-vehicle points are untouched and come first, in input order
    keep = ~np.isin(plain[:, 3], veh)
    assert np.array_equal(out[: keep.sum()], plain[keep])


# the three launch files that switch the filter on (ros/launch/odometry.launch.py:50, odometry_360.launch.py:50, odometry_raw.launch.py)
DYNAMIC_VARIANTS = {
    "odometry": (dict(), 40),
    "odometry_360": (dict(voxel_size_map=1.0, sem_th=0.8), 12),
    "odometry_raw": (dict(sem_th=0.2), 12),
}


@pytest.mark.parametrize("variant", list(DYNAMIC_VARIANTS))
def test_register_frame_with_dynamic_vehicle_filter(orc, variant):
    """Full-size (64 x 1875 rays) drive with dynamic_vehicle_filter = true against the oracle emitting the kept vehicle points in
    the reference's cluster order: every frame within the north-star tolerance, identical down-sampled clouds."""
    import sage_icp_b200 as sg
    from sage_icp_b200 import synthetic as syn
    from sage_icp_b200.config import launch_config
    overrides, frames = DYNAMIC_VARIANTS[variant]
    cfg = launch_config(dynamic_vehicle_filter=True, **overrides)
    gp, op = sg.SagePipeline(cfg), orc.OraclePipeline(cfg, threads=orc.max_threads(), evict_faithful=False)
    op.set_dynamic_cluster_order(True)
    traj = syn.trajectory(frames)
    worst = (0.0, 0.0)
    for i in range(frames):
        scan = syn.make_scan(700 + i, tuple(traj[i]))
        pg, _, _ = gp.register_frame(scan)
        po, _, _ = op.register_frame(scan)
        dt, da = pose_delta(pg, po)
        worst = (max(worst[0], dt), max(worst[1], da))
        assert dt <= POSE_TOL_M and da <= POSE_TOL_RAD, (variant, i, dt, da)
        assert np.array_equal(gp.last_source(), op.last_source()), (variant, i)
        assert np.array_equal(gp.last_frame_downsample(), op.last_frame_downsample()), (variant, i)
    assert worst[0] < 1e-6, worst  # observed: rounding level


def _pointcloud2_buffer(scan, label_f32=False):
    """The reference publishers' wire format (eval/kitti_pub.py:184-207): packed structured records."""
    fields = [("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("label", "<f4" if label_f32 else "u1"), ("rgb", "<u4")]
    rec = np.zeros(len(scan), dtype=np.dtype(fields))  # packed: itemsize 17 (u8 label) / 20 (f32 label)
    rec["x"], rec["y"], rec["z"] = scan[:, 0], scan[:, 1], scan[:, 2]
    rec["label"] = scan[:, 3]
    rec["rgb"] = 0x00ff00
    return rec.view(np.uint8).reshape(-1), rec.dtype.itemsize


@pytest.mark.parametrize("label_f32", [False, True])
def test_register_frame_from_pointcloud2_buffer(orc, cfg, label_f32):
    """utils::PointCloud2ToEigen (ros/ros2/Utils.hpp:161-180) done on the device: feeding the raw message buffer gives the same
    poses, bit for bit, as widening on the host and calling RegisterFrame(points); and both match the oracle."""
    import sage_icp_b200 as sg
    from sage_icp_b200 import synthetic as syn
    a, b, o = sg.SagePipeline(cfg), sg.SagePipeline(cfg), orc.OraclePipeline(cfg, evict_faithful=False)
    traj = syn.trajectory(6)
    for i in range(6):
        scan = syn.make_scan(500 + i, tuple(traj[i]), n_beams=32, n_az=700)
        buf, step = _pointcloud2_buffer(scan, label_f32)
        assert step == (20 if label_f32 else 17)
        pa, _, _ = a.register_frame_pointcloud2(buf, len(scan), step, (0, 4, 8, 12), 7 if label_f32 else 2)
        pb, _, _ = b.register_frame(scan)
        po, _, _ = o.register_frame(scan)
        assert np.array_equal(pa, pb), i
        assert np.array_equal(a.last_source(), b.last_source())
        dt, da = pose_delta(pa, po)
        assert dt <= POSE_TOL_M and da <= POSE_TOL_RAD
    with pytest.raises(sg.SageError):
        a.register_frame_pointcloud2(buf, len(scan), 10, (0, 4, 8, 12), 2)  # offsets do not fit point_step


def test_empty_and_tiny_frames_mid_sequence(orc, cfg):
    """An empty scan, and one whose points are all cropped away, in the middle of a drive: no correspondences, the 6x6
    system is zero, the step is the identity and the pose is the constant-velocity guess — on both sides."""
    import sage_icp_b200 as sg
    from sage_icp_b200 import synthetic as syn
    gp, op = sg.SagePipeline(cfg), orc.OraclePipeline(cfg, evict_faithful=False)
    traj = syn.trajectory(6)
    frames = [syn.make_scan(900 + i, tuple(traj[i]), n_beams=32, n_az=600) for i in range(6)]
    frames[3] = np.zeros((0, 4))
    frames[4] = np.array([[0.5, 0.2, 0.1, 40.0], [300.0, 1.0, 0.0, 50.0], [1.0, -1.0, 0.3, 81.0]])  # all outside (min_range, max_range)
    for i, f in enumerate(frames):
        pg, _, _ = gp.register_frame(f)
        po, _, _ = op.register_frame(f)
        dt, da = pose_delta(pg, po)
        assert dt <= POSE_TOL_M and da <= POSE_TOL_RAD, (i, dt, da)
        assert gp.last_iterations() == op.last_iterations(), i
        assert len(gp.last_source()) == len(op.last_source())
    assert len(gp.poses()) == 6
