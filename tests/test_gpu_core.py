"""GPU parity tests of the core level (VoxelHashMap + Registration) through the C ABI, against the oracle."""
import numpy as np
import pytest

from conftest import POSE_TOL_M, POSE_TOL_RAD, assert_maps_equal, pose_delta

pytestmark = pytest.mark.gpu

BASIC_LABELS = [40, 44, 48, 49, 50, 70, 72]
IDENT = np.array([0, 0, 0, 0, 0, 0, 1.0])


def _maps(orc, voxel_size=0.8, max_distance=100.0, basic=20, critical=20, faithful=False):
    import sage_icp_b200 as sg
    g = sg.SageMap(voxel_size, max_distance, basic, critical, BASIC_LABELS)
    o = orc.OracleMap(voxel_size, max_distance, basic, critical, BASIC_LABELS, evict_faithful=faithful)
    return g, o


def _street_points(n, seed, x_lo=-60.0, x_hi=60.0):
    from sage_icp_b200 import synthetic as syn
    return syn.sample_street_map(n, seed, x_lo, x_hi)


def test_library_reports_device():
    import sage_icp_b200 as sg
    assert sg.device_count() >= 1


@pytest.mark.parametrize("basic,critical", [(20, 20), (3, 2), (1, 0), (5, 0)])
def test_add_points_matches_oracle(orc, basic, critical):
    """VoxelHashMap::AddPoints + VoxelBlock::AddPoint rule table: identical voxels, identical stored order."""
    g, o = _maps(orc, basic=basic, critical=critical)
    rng = np.random.default_rng(1)
    # dense cloud in a few voxels so every branch of AddPoint fires (append / drop / overwrite label-0 / critical append)
    pts = np.empty((6000, 4))
    pts[:, :3] = rng.uniform(-2.0, 2.0, (6000, 3))
    pts[:, 3] = rng.choice([0, 0, 40, 50, 70, 80, 81, 10, 252], 6000)
    for chunk in np.array_split(pts, 3):
        g.add_points(chunk)
        o.add_points(chunk)
        assert_maps_equal(g.dump(), o.dump())
    assert g.num_voxels() == o.num_voxels()
    assert g.num_points() == o.num_points()


def test_add_points_street_scale(orc):
    g, o = _maps(orc)
    pts = _street_points(400_000, 3)
    g.add_points(pts)
    o.add_points(pts)
    assert g.num_voxels() == o.num_voxels() and g.num_points() == o.num_points()
    assert_maps_equal(g.dump(), o.dump())
    # set-equality of the exported cloud (order is block order on the device, robin order in the reference)
    a, b = g.pointcloud(), o.pointcloud()
    assert a.shape == b.shape
    assert np.array_equal(a[np.lexsort(a.T[::-1])], b[np.lexsort(b.T[::-1])])


def test_update_and_evict_clean_mode(orc):
    """Update(points, pose) = transform + AddPoints + RemovePointsFarFromLocation (clean eviction on both sides)."""
    g, o = _maps(orc, max_distance=30.0)
    rng = np.random.default_rng(5)
    for f in range(12):
        pose = orc.se3_exp([4.0 * f, 0.3 * np.sin(f), 0.0, 0.0, 0.0, 0.02 * f])
        local = np.empty((5000, 4))
        local[:, :3] = rng.uniform(-25, 25, (5000, 3)) * np.array([1, 1, 0.1])
        local[:, 3] = rng.choice([0, 40, 50, 80], 5000)
        g.update(local, pose)
        o.update(local, pose)
        assert g.num_voxels() == o.num_voxels(), f"frame {f}"
        assert_maps_equal(g.dump(), o.dump())
    assert o.num_voxels() > 0


def test_update_and_evict_faithful_mode(orc):
    """sage_map_set_eviction(m, 1): the reference's erase-while-iterating sweep over its tsl::robin_map, bucket for bucket
    (core/VoxelHashMap.cpp:176-184 leaves some far voxels behind), and the map's iteration order — the dump (voxel order,
    stored points) and the exported cloud equal the oracle's in ORDER, frame after frame; far survivors really occur; blocks
    freed by the sweep are reused; load() and clear() keep the mirror in step."""
    g, o = _maps(orc, max_distance=30.0, faithful=True)
    g.set_eviction(True)
    clean, _ = _maps(orc, max_distance=30.0)
    rng = np.random.default_rng(5)
    leftovers = 0
    for f in range(12):
        pose = orc.se3_exp([4.0 * f, 0.3 * np.sin(f), 0.0, 0.0, 0.0, 0.02 * f])
        local = np.empty((5000, 4))
        local[:, :3] = rng.uniform(-25, 25, (5000, 3)) * np.array([1, 1, 0.1])
        local[:, 3] = rng.choice([0, 40, 50, 80], 5000)
        g.update(local, pose)
        o.update(local, pose)
        clean.update(local, pose)
        gk, gc, gp = g.dump()
        ok, oc, op = o.dump()
        assert np.array_equal(gk, ok), f"frame {f}: voxel order"
        assert np.array_equal(gc, oc) and np.array_equal(gp, op), f"frame {f}"
        assert np.array_equal(g.pointcloud(), o.pointcloud()), f"frame {f}: Pointcloud() order"
        leftovers += g.num_voxels() - clean.num_voxels()
    assert leftovers > 0  # the faithful sweep kept voxels the clean one dropped
    # queries see the left-over far voxels exactly as the reference would
    q = np.c_[rng.uniform(-60, 60, (4000, 2)) + [44.0, 0.0], rng.uniform(-2, 2, 4000), rng.choice([0, 40, 50, 80], 4000).astype(float)]
    tg, mg = g.get_correspondences(q, 2.0, 0.4)
    src, tgt, qi = o.get_correspondences(q, 2.0, 0.4)
    assert np.array_equal(np.flatnonzero(mg), qi) and np.array_equal(tg[mg], tgt)
    # a snapshot loaded into a faithful map takes the snapshot's order as insertion order, like re-inserting it would
    g2, o2 = _maps(orc, max_distance=30.0, faithful=True)
    g2.set_eviction(True)
    g2.load(*o.dump())
    for k, c, p in zip(*o.dump()):
        o2.add_points(p[:c])
    assert np.array_equal(g2.dump()[0], o2.dump()[0])
    import sage_icp_b200 as sg
    with pytest.raises(sg.SageError):
        g.set_eviction(False)  # only on an empty map
    g.clear(); o.clear()
    g.update(local, pose); o.update(local, pose)
    assert np.array_equal(g.dump()[0], o.dump()[0])  # clear() keeps the bucket count, on both sides


def test_empty_inputs(orc):
    g, o = _maps(orc)
    assert g.empty() and g.num_voxels() == 0 and g.num_points() == 0
    g.add_points(np.zeros((0, 4)))
    assert g.empty()
    pose, it = g.register_frame(np.ones((10, 4)), IDENT, 1.0, 0.3, 0.4)
    assert it == 0 and np.array_equal(pose, IDENT)  # empty map -> initial guess (core/Registration.cpp:119)
    tgt, matched = g.get_correspondences(np.ones((10, 4)), 1.0, 0.4)
    assert not matched.any()
    g.add_points(np.array([[0.1, 0.1, 0.1, 40.0]]))
    assert not g.empty() and g.num_voxels() == 1
    g.clear()
    assert g.empty()


def _query_cloud(seed, n, spread=40.0):
    from sage_icp_b200 import synthetic as syn
    scan = syn.make_scan(seed, (0.0, 0.0, 0.0), n_beams=32, n_az=max(8, n // 32))
    pose = syn.pose7_from_xyyaw((0.3, -0.2, 0.01))
    r = np.linalg.norm(scan[:, :3], axis=1)
    scan = scan[(r > 3) & (r < 60)]
    return scan, pose


@pytest.mark.parametrize("sem_th", [0.4, 0.05, 1.0])
def test_correspondences_bit_exact(orc, sem_th):
    """GetCorrespondences: per query the same target point (bit-exact) and the same accept decision as the oracle,
    including queries whose 27-neighbourhood is empty and label-0 / unknown-label queries."""
    from oracle.oracle_py import se3_act
    g, o = _maps(orc)
    pts = _street_points(300_000, 7)
    o.add_points(pts)
    g.load(*o.dump())
    assert_maps_equal(g.dump(), o.dump())
    scan, pose = _query_cloud(11, 16000)
    q = scan.copy()
    from sage_icp_b200.synthetic import SENSOR_HEIGHT
    q[:, 2] += SENSOR_HEIGHT
    q[:, 0] += 0.3
    # a few far-away queries with empty neighbourhoods
    q[:50, :3] += 500.0
    max_dist = 1.5
    tgt, matched = g.get_correspondences(q, max_dist, sem_th)
    src_o, tgt_o, qidx = o.get_correspondences(q, max_dist, sem_th)
    m_o = np.zeros(len(q), bool)
    m_o[qidx] = True
    assert matched.sum() > 1000
    assert np.array_equal(matched, m_o)
    assert np.array_equal(tgt[matched], tgt_o)
    # independent brute-force check of the oracle itself on a subset (numpy, all map points)
    allp = o.pointcloud()
    for i in np.flatnonzero(matched)[:40]:
        d = ((allp[:, :3] - q[i, :3]) ** 2).sum(1)
        same = (allp[:, 3].astype(int) == int(q[i, 3])) | ((allp[:, 3] * q[i, 3]).astype(int) == 0)
        m = np.where(same, d * sem_th, d)
        # restrict to the 27-voxel neighbourhood as the reference does
        k = (allp[:, :3] / 0.8).astype(int) - (q[i, :3] / 0.8).astype(int)
        m[np.abs(k).max(1) > 1] = np.inf
        assert np.isclose(m.min(), (np.where((int(tgt[i, 3]) == int(q[i, 3])) or int(tgt[i, 3] * q[i, 3]) == 0, sem_th, 1.0)
                                    * ((tgt[i, :3] - q[i, :3]) ** 2).sum()), rtol=1e-12)


def test_normal_equations_match_oracle(orc):
    g, o = _maps(orc)
    o.add_points(_street_points(300_000, 7))
    g.load(*o.dump())
    scan, _ = _query_cloud(13, 16000)
    from sage_icp_b200.synthetic import SENSOR_HEIGHT
    q = scan.copy()
    q[:, 2] += SENSOR_HEIGHT
    q[:, 1] += 0.2
    kernel = 0.5
    JTJ, JTr, n = g.normal_equations(q, 1.5, kernel, 0.4)
    src_o, tgt_o, _ = o.get_correspondences(q, 1.5, 0.4)
    JTJ_o, JTr_o, x_o, est_o = orc.align_clouds(src_o, tgt_o, kernel)
    assert n == len(src_o)
    assert np.allclose(JTJ, JTJ_o, rtol=1e-11, atol=1e-9 * np.abs(JTJ_o).max())
    assert np.allclose(JTr, JTr_o, rtol=1e-10, atol=1e-9 * np.abs(JTr_o).max())


@pytest.mark.parametrize("n_queries,guess_err", [(4000, (0.3, 0.1, 0.01)), (30000, (0.25, -0.2, -0.015)), (33, (0.1, 0.0, 0.0))])
def test_core_register_frame_pose_parity(orc, n_queries, guess_err):
    """sage_icp::RegisterFrame: same iteration count, pose within 1e-4 m / 1e-5 rad of the oracle."""
    from sage_icp_b200 import synthetic as syn
    g, o = _maps(orc)
    o.add_points(_street_points(600_000, 21, -80.0, 80.0))
    g.load(*o.dump())
    scan = syn.make_scan(3, (0.0, 0.0, 0.0), n_beams=64, n_az=max(2, n_queries // 40))
    r = np.linalg.norm(scan[:, :3], axis=1)
    scan = scan[(r > 5) & (r < 70)][:n_queries]
    guess = syn.pose7_from_xyyaw(guess_err)
    pose_o, it_o = o.register_frame_core(scan, guess, 3.0, 0.33, 0.4)
    pose_g, it_g = g.register_frame(scan, guess, 3.0, 0.33, 0.4)
    dt, da = pose_delta(pose_g, pose_o)
    assert it_g == it_o, (it_g, it_o)
    assert dt <= POSE_TOL_M and da <= POSE_TOL_RAD, (dt, da)
    assert dt < 1e-9 and da < 1e-10  # in practice parity is at rounding level
    # fixed-iteration mode used by the kernel-level bench (10 GN iterations, no early exit)
    pose_o, it_o = o.register_frame_core(scan, guess, 3.0, 0.33, 0.4, max_iters=10, est_th=0.0)
    pose_g, it_g = g.register_frame(scan, guess, 3.0, 0.33, 0.4, max_iters=10, est_th=0.0)
    assert it_g == it_o == 10
    dt, da = pose_delta(pose_g, pose_o)
    assert dt <= POSE_TOL_M and da <= POSE_TOL_RAD, (dt, da)


def test_register_frame_device_resident_matches_host_path(orc):
    import torch
    from sage_icp_b200 import synthetic as syn
    g, o = _maps(orc)
    o.add_points(_street_points(300_000, 23))
    g.load(*o.dump())
    scan = syn.make_scan(4, (0.0, 0.0, 0.0), n_beams=32, n_az=300)
    guess = syn.pose7_from_xyyaw((0.2, 0.1, 0.0))
    pose_h, it_h = g.register_frame(scan, guess, 3.0, 0.33, 0.4)
    d = torch.from_numpy(scan).cuda()
    torch.cuda.synchronize()
    pose_d, it_d = g.register_frame_device(d.data_ptr(), len(scan), guess, 3.0, 0.33, 0.4)
    assert it_h == it_d and np.array_equal(pose_h, pose_d)
    assert torch.equal(d.cpu(), torch.from_numpy(scan))  # the caller's frame is not modified (the reference copies it)


@pytest.mark.parametrize("n_az", [60, 300])
def test_persistent_loop_equals_one_launch_per_iteration(orc, monkeypatch, n_az):
    """Small scans run their whole Gauss-Newton loop in one cooperative launch (nn_search_persistent_kernel); the result is
    bit-identical to one launch per iteration (SAGE_PERSISTENT_MAX=0 switches the former off), for the warp-per-query mode
    (60 x 32 queries) and the thread-per-query mode (300 x 32), and both equal the oracle within tolerance."""
    import sage_icp_b200 as sg
    from sage_icp_b200 import synthetic as syn
    o = orc.OracleMap(0.8, 100.0, 20, 20, BASIC_LABELS, evict_faithful=False)
    o.add_points(_street_points(300_000, 29))
    scan = syn.make_scan(6, (0.0, 0.0, 0.0), n_beams=32, n_az=n_az)
    guess = syn.pose7_from_xyyaw((0.25, -0.1, 0.01))
    out = []
    for persistent_max in ("0", "20000"):
        monkeypatch.setenv("SAGE_PERSISTENT_MAX", persistent_max)  # read when the map's search is first configured
        g = sg.SageMap(0.8, 100.0, 20, 20, BASIC_LABELS)
        g.load(*o.dump())
        before = sg.launch_count()
        pose, it = g.register_frame(scan, guess, 3.0, 0.33, 0.4)
        out.append((pose, it, sg.launch_count() - before))
        pose_again, it_again = g.register_frame(scan, guess, 3.0, 0.33, 0.4)
        assert it_again == it and np.array_equal(pose_again, pose)  # reproducible
    (p0, it0, l0), (p1, it1, l1) = out
    assert it0 == it1 and it0 > 2 and np.array_equal(p0, p1)
    assert l1 < l0 and l1 <= 3  # init + one cooperative launch (+ nothing per iteration)
    po, ito = o.register_frame_core(scan, guess, 3.0, 0.33, 0.4)
    dt, da = pose_delta(p1, po)
    assert ito == it1 and dt <= POSE_TOL_M and da <= POSE_TOL_RAD


def test_nn_stats_match_oracle(orc):
    g, o = _maps(orc)
    o.add_points(_street_points(200_000, 9))
    g.load(*o.dump())
    scan, _ = _query_cloud(17, 8000)
    assert g.nn_stats(scan) == o.nn_stats(scan)


def test_degenerate_batch_into_one_voxel(orc):
    """150 000 points of one batch in a single voxel (and 50 000 spread around): the per-voxel replay must stay in input order
    and finish promptly (the arrival list is merge-sorted, not selected quadratically)."""
    import time
    g, o = _maps(orc)
    rng = np.random.default_rng(8)
    a = np.c_[rng.uniform(0.01, 0.79, (150_000, 3)), rng.choice([0, 40, 81], 150_000)]
    b = np.c_[rng.uniform(-20, 20, (50_000, 3)), rng.choice([0, 40, 81], 50_000)]
    pts = np.concatenate([a, b])
    rng.shuffle(pts)
    t = time.time()
    g.add_points(pts)
    assert g.num_voxels() > 1000
    assert time.time() - t < 20.0
    o.add_points(pts)
    assert_maps_equal(g.dump(), o.dump())


@pytest.mark.parametrize("case", ["one_pair", "two_pairs", "collinear"])
def test_rank_deficient_systems_take_a_bounded_step(orc, case):
    """One or two correspondences, or collinear ones (track loss, a scan leaving the map): the 3x3 rotation block of the normal
    equations is singular up to rounding.  Dividing by that noise would throw the pose far away; the step must degrade to the
    translation that aligns the weighted centroids instead (registration.cu icp_solve_xi).  The query ends on its target, the
    rotation stays the identity, nothing explodes."""
    import sage_icp_b200 as sg
    g = sg.SageMap(0.8, 1e9, 20, 20, BASIC_LABELS)
    if case == "one_pair":
        pts = np.array([[10.3, 4.1, 0.7, 40.0]])
    elif case == "two_pairs":
        pts = np.array([[10.3, 4.1, 0.7, 40.0], [30.9, -7.2, 1.1, 50.0]])
    else:
        t = np.linspace(0.0, 40.0, 30)
        pts = np.c_[5.0 + t, 2.0 + 0.5 * t, 0.3 + 0.1 * t, np.full(30, 40.0)]
    g.add_points(pts)
    q = pts.copy()
    q[:, :3] += np.array([0.21, -0.13, 0.08])  # a pure translation well inside max_dist
    guess = np.array([0, 0, 0, 0, 0, 0, 1.0])
    pose, it = g.register_frame(q, guess, 3.0, 1.0 / 3.0, 0.4, max_iters=50)
    assert np.all(np.isfinite(pose)) and it <= 50
    assert np.linalg.norm(pose[:3] - np.array([-0.21, 0.13, -0.08])) < 1e-6, pose
    assert np.linalg.norm(pose[3:6]) < 1e-9 and abs(pose[6] - 1.0) < 1e-12, pose
