"""BASELINE configs[2] in miniature: a KITTI-shaped synthetic drive with the map built incrementally ON THE DEVICE
(VoxelHashMap::Update every frame, eviction and block reuse included), frame-by-frame against the oracle."""
import numpy as np
import pytest

from conftest import POSE_TOL_M, POSE_TOL_RAD, map_as_dict, pose_delta

pytestmark = pytest.mark.gpu


def test_streaming_drive_parity_with_eviction(orc):
    import sage_icp_b200 as sg
    from sage_icp_b200 import synthetic as syn
    from sage_icp_b200.config import launch_config
    cfg = launch_config(local_map_range=40.0)  # short map horizon: voxels are evicted and their blocks reused
    gp, op = sg.SagePipeline(cfg), orc.OraclePipeline(cfg, threads=orc.max_threads(), evict_faithful=False)
    n = 100
    traj = syn.trajectory(n)
    worst_t = worst_r = 0.0
    peak_voxels = 0
    for i in range(n):
        scan = syn.make_scan(1000 + i, tuple(traj[i]), n_beams=64, n_az=1875)
        pg, t_icp, t_all = gp.register_frame(scan)
        po, _, _ = op.register_frame(scan)
        dt, da = pose_delta(pg, po)
        worst_t, worst_r = max(worst_t, dt), max(worst_r, da)
        assert dt <= POSE_TOL_M and da <= POSE_TOL_RAD, (i, dt, da)
        assert gp.last_iterations() == op.last_iterations(), i
        if i % 10 == 0:
            assert np.array_equal(gp.last_source(), op.last_source()), i
            nv = gp.map().num_voxels()
            assert nv == op.map().num_voxels(), i
            peak_voxels = max(peak_voxels, nv)
    # the drive covered ~95 m with a 40 m horizon: the map stopped growing because early voxels were evicted
    assert gp.map().num_voxels() == op.map().num_voxels() < peak_voxels
    assert worst_t < 1e-7 and worst_r < 1e-8, (worst_t, worst_r)  # parity is at rounding level, not merely within tolerance
    final_t, final_r = pose_delta(gp.poses()[-1], op.poses()[-1])
    assert final_t <= POSE_TOL_M and final_r <= POSE_TOL_RAD
    # identical maps at the end (same voxels, same stored order; points agree to the poses' rounding)
    dg, do = map_as_dict(*gp.map().dump()), map_as_dict(*op.map().dump())
    assert set(dg) == set(do)
    for k in dg:
        assert dg[k].shape == do[k].shape and np.allclose(dg[k], do[k], atol=1e-7, rtol=0), k
    # the vehicle really moved and the estimate followed it
    assert gp.poses()[-1][0] > 0.8 * (traj[-1][0] - traj[0][0])


def test_deskewed_drive_parity(orc):
    """RegisterFrame(frame, timestamps) with config.deskew = true: DeSkewScan on the device (core/Deskew.cpp:36-50)."""
    import sage_icp_b200 as sg
    from sage_icp_b200 import synthetic as syn
    from sage_icp_b200.config import launch_config
    cfg = launch_config(deskew=True)
    gp, op = sg.SagePipeline(cfg), orc.OraclePipeline(cfg, evict_faithful=False)
    traj = syn.trajectory(12)
    for i in range(12):
        scan = syn.make_scan(2000 + i, tuple(traj[i]), n_beams=32, n_az=600)
        ts = (np.arange(len(scan)) % 600) / 600.0  # azimuth sweep: 0..1 over one revolution, per beam
        pg, _, _ = gp.register_frame(scan, ts)
        po, _, _ = op.register_frame(scan, ts)
        dt, da = pose_delta(pg, po)
        assert dt <= POSE_TOL_M and da <= POSE_TOL_RAD, (i, dt, da)
        assert np.allclose(gp.last_source(), op.last_source(), atol=1e-9), i  # de-skewed points differ at rounding level only
    assert len(gp.poses()) == 12


def test_streaming_drive_with_the_reference_eviction_quirk(orc):
    """The same kind of drive with sage_map_set_eviction(map, 1) against the oracle in ITS faithful mode: the far voxels the
    reference's erase-while-iterating sweep leaves behind stay searchable on both sides, so poses, iteration counts and the
    map agree frame by frame — and LocalMap() comes out in the reference's order (tsl::robin_map iteration order)."""
    import sage_icp_b200 as sg
    from sage_icp_b200 import synthetic as syn
    from sage_icp_b200.config import launch_config
    cfg = launch_config(local_map_range=30.0)
    gp, op = sg.SagePipeline(cfg), orc.OraclePipeline(cfg, threads=orc.max_threads(), evict_faithful=True)
    gp.map().set_eviction(True)
    clean = sg.SagePipeline(cfg)
    n = 60
    traj = syn.trajectory(n)
    extra = 0
    for i in range(n):
        scan = syn.make_scan(4000 + i, tuple(traj[i]), n_beams=32, n_az=1200)
        pg, _, _ = gp.register_frame(scan)
        po, _, _ = op.register_frame(scan)
        clean.register_frame(scan)
        dt, da = pose_delta(pg, po)
        assert dt <= POSE_TOL_M and da <= POSE_TOL_RAD, (i, dt, da)
        assert gp.last_iterations() == op.last_iterations(), i
        if i % 6 == 5:
            gk, gc, gpts = gp.map().dump()
            ok, oc, opts = op.map().dump()
            assert np.array_equal(gk, ok) and np.array_equal(gc, oc), i  # same voxels in the same (robin) order
            assert np.allclose(gpts, opts, atol=1e-7, rtol=0)
            extra = max(extra, gp.map().num_voxels() - clean.map().num_voxels())
    assert extra > 0  # the quirk was exercised: voxels a clean sweep would have dropped are still in the map
    a, b = gp.local_map(), op.local_map()
    assert a.shape == b.shape and np.array_equal(a[:, 3], b[:, 3]) and np.allclose(a, b, atol=1e-7, rtol=0)
    gp.reinitialize(); op.reset()
    assert gp.map().empty()
    scan = syn.make_scan(4000, tuple(traj[0]), n_beams=32, n_az=1200)
    gp.register_frame(scan); op.register_frame(scan)
    assert np.array_equal(gp.map().dump()[0], op.map().dump()[0])  # Clear() kept the bucket count on both sides
