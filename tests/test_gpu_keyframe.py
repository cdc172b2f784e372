"""The ROS node's key-frame test on the device (ros/ros2/OdometryServer.cpp:222-241): utils::EigenToGridMap of the scan moved into
the last key frame (sageICP::TransformToLastFrame) and utils::compute_occ_overlap against the last key frame's grid
(ros/ros2/Utils.hpp:220-258) — bit-identical grids, equal overlap; and the bulk pose read-out the node's per-scan poses() needs."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BOUNDS = [[-51.2, 51.2], [-51.2, 51.2], [-4.0, 2.4]]  # ros/launch/odometry.launch.py:88
H, W = 128, 128                                        # :90


def test_key_frame_grid_matches_the_oracle(orc, cfg):
    import sage_icp_b200 as sg
    from sage_icp_b200 import synthetic as syn
    p = sg.SagePipeline(cfg)
    scan = syn.make_scan(5, (0.0, 0.0, 0.0))
    g, ov = p.key_frame_grid(scan, BOUNDS, H, W)
    assert ov is None
    go = orc.grid_map(scan, BOUNDS, H, W)
    assert np.array_equal(g, go) and 1000 < g.sum() < H * W
    # the next scan, moved into the first one's frame, against the first grid
    last, cur = orc.se3_exp([0, 0, 0, 0, 0, 0]), orc.se3_exp([2.5, 0.2, 0.01, 0.001, -0.002, 0.03])
    scan2 = syn.make_scan(6, (2.5, 0.2, 0.03))
    g2, ov2 = p.key_frame_grid(scan2, BOUNDS, H, W, last_pose=last, current_pose=cur, last_occ=g)
    moved = p.transform_to_last_frame(last, cur, scan2)
    go2 = orc.grid_map(moved, BOUNDS, H, W)
    assert np.array_equal(g2, go2)
    assert ov2 == orc.occ_overlap(go, go2) and 0.3 < ov2 < 1.0
    # non-square grid, asymmetric bounds (the reference adds the UPPER bound before dividing), points on the bounds, NaN, empty input
    b = [[-20.0, 60.0], [-10.0, 30.0], [-1.0, 5.0]]
    pts = scan.copy()
    pts[:10, 0], pts[10:20, 1], pts[20:30, 2] = 60.0, -10.0, np.nan
    g3, _ = p.key_frame_grid(pts, b, 40, 96)
    assert np.array_equal(g3, orc.grid_map(pts, b, 40, 96))
    g4, ov4 = p.key_frame_grid(np.zeros((0, 4)), BOUNDS, H, W, last_occ=np.zeros((H, W), np.int32))
    assert g4.sum() == 0 and np.isnan(ov4)  # 0 / 0, as utils::compute_occ_overlap
    with pytest.raises(sg.SageError):
        p.key_frame_grid(scan, BOUNDS, 0, W)


def test_poses_bulk_readout(orc, cfg):
    import sage_icp_b200 as sg
    from sage_icp_b200 import synthetic as syn
    p = sg.SagePipeline(cfg)
    traj = syn.trajectory(5)
    got = []
    for i in range(5):
        pose, _, _ = p.register_frame(syn.make_scan(40 + i, tuple(traj[i]), n_beams=32, n_az=500))
        got.append(pose)
        assert np.array_equal(p.poses(first=i), np.array(got[i:]))  # the node's pattern: only the new tail
    assert np.array_equal(p.poses(), np.array(got))
    assert np.array_equal(p.pose(3), got[3]) and len(p.poses(first=5)) == 0
    with pytest.raises(sg.SageError):
        p.poses(first=6)
    p.reinitialize()
    assert len(p.poses()) == 0
