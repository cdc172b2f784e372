"""Oracle pinning, part 4: the oracle's restatement against the REFERENCE'S OWN CODE run here.

oracle/_ref/libsage_ref.so is the reference's unmodified hot-path sources (cpp/sage_icp/core/{Deskew,Preprocessing,
Registration,Threshold,VoxelHashMap}.cpp, pipeline/sageICP.cpp), compiled where they lie under /root/reference against stand-in
headers for the third-party libraries this image lacks (oracle/shim/: Eigen, Sophus, oneTBB, tsl::robin_map, PCL), with a thin C
wrapper (oracle/ref_capi.cpp).  What these tests pin is the reference's own logic — crop rule, group lookup, truncation keys,
first-point-per-voxel, the AddPoint table, the 27-voxel scan with the semantic metric and strict '<', the acceptance test, the
Jacobian / weights / normal equations, the ICP loop and its stopping rule, the erase-while-iterating sweep, the adaptive threshold
state machine, the pipeline's call order.  The third-party arithmetic (pivoted LDLT, SE3 exp/log, robin_map bucket order, PCL
cluster order) is the stand-ins' restatement of the same published formulas the oracle uses, so it is NOT pinned by this file
(tests/test_oracle_se3.py, test_oracle_robin.py do that against scipy / a Python model).

The library is built by oracle/Makefile (target `ref`) wherever /root/reference exists and travels to the GPU box with the
snapshot; where neither the library nor the reference is present the tests skip."""
import numpy as np
import pytest

from conftest import map_as_dict, pose_delta

BASIC_LABELS = [40, 44, 48, 49, 50, 70, 72]


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_py
    if not ref_py.available():
        pytest.skip("neither oracle/_ref/libsage_ref.so nor /root/reference is present")
    ref_py.lib()
    return ref_py


def _scan(seed, pose=(0.0, 0.0, 0.0), beams=32, az=600):
    from sage_icp_b200 import synthetic as syn
    return syn.make_scan(seed, pose, n_beams=beams, n_az=az)


def test_reference_build_is_the_reference(ref):
    """The library was made from the reference's sources, not from anything in this repository: its recipe names them."""
    import os
    mk = open(os.path.join(os.path.dirname(ref.__file__), "Makefile")).read()
    for src in ("core/Deskew.cpp", "core/Preprocessing.cpp", "core/Registration.cpp", "core/Threshold.cpp", "core/VoxelHashMap.cpp",
                "pipeline/sageICP.cpp"):
        assert f"$(REF)/sage_icp/{src}" in mk
    assert "REF ?= /root/reference/cpp" in mk


def test_preprocess_range_branch(ref, orc, cfg):
    scan = _scan(1)
    scan[:50, :3] *= 0.01  # inside min_range
    scan[50:100, :3] *= 40.0  # beyond max_range
    a, b = ref.preprocess(cfg, scan), orc.preprocess(scan, cfg.max_range, cfg.min_range, cfg.label_max_range)
    assert a.shape == b.shape and np.array_equal(a, b)
    assert len(a) < len(scan) and (a[:, 3] == 0).sum() > (scan[:, 3] == 0).sum()


@pytest.mark.parametrize("scale", [0.5, 1.5, 1.0])
@pytest.mark.parametrize("beams,az", [(32, 600), (64, 1200), (4, 30)])
def test_voxel_downsample_including_order(ref, orc, cfg, scale, beams, az):
    """Same survivors in the same order: group lookup, dropped labels, truncation keys, first point per voxel, and the iteration
    order of the (unreserved) per-group maps."""
    pts = orc.preprocess(_scan(2, beams=beams, az=az), cfg.max_range, cfg.min_range, cfg.label_max_range)
    a, b = ref.voxel_downsample(cfg, pts, scale), orc.voxel_downsample(cfg, pts, scale)
    assert a.shape == b.shape and np.array_equal(a, b)


def test_voxel_downsample_negative_and_fractional_inputs(ref, orc, cfg):
    rng = np.random.default_rng(3)
    pts = np.c_[rng.uniform(-30, 30, (8000, 3)), rng.choice([0, 10, 40, 48, 50, 70, 80, 81, 99, 30, 252, 40.7, 0.4], 8000)]
    for scale in (0.5, 1.5):
        assert np.array_equal(ref.voxel_downsample(cfg, pts, scale), orc.voxel_downsample(cfg, pts, scale))


@pytest.mark.parametrize("basic,critical", [(20, 20), (3, 2), (1, 0), (5, 0)])
def test_add_points_rule_table(ref, orc, basic, critical):
    """VoxelBlock::AddPoint, every branch (append / drop label 0 / basic label overwrites a stored label-0 point / critical label
    appends up to basic+critical, then overwrites): identical voxels in identical map order with identical stored order."""
    r = ref.RefMap(0.8, 100.0, basic, critical, BASIC_LABELS)
    o = orc.OracleMap(0.8, 100.0, basic, critical, BASIC_LABELS, evict_faithful=True)
    rng = np.random.default_rng(1)
    pts = np.c_[rng.uniform(-2.0, 2.0, (6000, 3)), rng.choice([0, 0, 40, 50, 70, 80, 81, 10, 252], 6000)]
    for chunk in np.array_split(pts, 3):
        r.add_points(chunk)
        o.add_points(chunk)
        (rk, rc, rp), (ok, oc, op) = r.dump(), o.dump()
        assert np.array_equal(rk, ok) and np.array_equal(rc, oc) and np.array_equal(rp, op)
    assert np.array_equal(r.pointcloud(), o.pointcloud())


def test_update_and_the_erase_while_iterating_sweep(ref, orc):
    """Update(points, pose) over a moving origin with a short horizon: transform, insert, RemovePointsFarFromLocation exactly as
    the reference does it (range-for over the map with erase inside).  The oracle's faithful mode reproduces the survivors and
    the order frame by frame; its clean mode drops strictly more."""
    r = ref.RefMap(0.8, 30.0, 20, 20, BASIC_LABELS)
    o = orc.OracleMap(0.8, 30.0, 20, 20, BASIC_LABELS, evict_faithful=True)
    clean = orc.OracleMap(0.8, 30.0, 20, 20, BASIC_LABELS, evict_faithful=False)
    rng = np.random.default_rng(5)
    extra = 0
    for f in range(12):
        pose = orc.se3_exp([4.0 * f, 0.3 * np.sin(f), 0.0, 0.0, 0.0, 0.02 * f])
        local = np.c_[rng.uniform(-25, 25, (5000, 3)) * [1, 1, 0.1], rng.choice([0, 40, 50, 80], 5000)]
        r.update(local, pose); o.update(local, pose); clean.update(local, pose)
        (rk, rc, rp), (ok, oc, op) = r.dump(), o.dump()
        assert np.array_equal(rk, ok), f
        assert np.array_equal(rc, oc) and np.array_equal(rp, op), f
        extra += r.num_voxels() - clean.num_voxels()
    assert extra > 0
    r.clear(); o.clear()
    assert r.empty()
    r.update(local, pose); o.update(local, pose)
    assert np.array_equal(r.dump()[0], o.dump()[0])  # Clear() keeps the bucket count on both sides


@pytest.mark.parametrize("sem_th", [0.4, 0.05, 1.0])
def test_get_correspondences(ref, orc, sem_th):
    """The 27-voxel scan in enumeration order with the semantic metric (equal labels or a zero label shrink the distance by th),
    strict '<' from DBL_MAX, acceptance on the TRUE distance; queries with an empty neighbourhood yield no pair on both sides
    (the reference reads an unset vector there — NaN in this build, see oracle/shim/Eigen/Core)."""
    from sage_icp_b200 import synthetic as syn
    pts = syn.sample_street_map(150_000, 7, -40.0, 40.0)
    r = ref.RefMap(0.8, 100.0, 20, 20, BASIC_LABELS)
    o = orc.OracleMap(0.8, 100.0, 20, 20, BASIC_LABELS)
    r.add_points(pts); o.add_points(pts)
    scan = _scan(11, beams=32, az=400)
    q = scan.copy()
    q[:, 2] += 1.73  # sensor height: into the map frame
    q[:, 0] += 0.3
    q = np.r_[q, [[500.0, 500.0, 50.0, 40.0]], [[-0.3, 0.2, 0.1, 0.0]]]  # one query far from everything, one at the origin voxel
    rs, rt = r.get_correspondences(q, 2.0, sem_th)
    os_, ot, oq = o.get_correspondences(q, 2.0, sem_th)
    assert len(rs) == len(os_) > 1000 and len(rs) < len(q)
    assert np.array_equal(rs, os_) and np.array_equal(rt, ot)
    assert np.array_equal(rs, q[oq])


def test_register_frame_core(ref, orc):
    """sage_icp::RegisterFrame: TransformPoints, GetCorrespondences, AlignClouds (Jacobian [I | -hat(s)], Geman-McClure weights,
    6x6 normal equations, LDLT, exp), T <- est * T, stop on |log(est)| < 1e-4 or 500 iterations.  Same pose to rounding (the
    reference accumulates J^T w J as 6x6 products, the oracle as 16 scalar sums: different summation trees)."""
    from sage_icp_b200 import synthetic as syn
    pts = syn.sample_street_map(200_000, 23, -50.0, 50.0)
    r = ref.RefMap(0.8, 100.0, 20, 20, BASIC_LABELS)
    o = orc.OracleMap(0.8, 100.0, 20, 20, BASIC_LABELS)
    r.add_points(pts); o.add_points(pts)
    for seed, g in ((4, (0.25, -0.1, 0.01)), (5, (-0.3, 0.2, -0.015)), (6, (0.0, 0.0, 0.0))):
        scan = _scan(seed, beams=32, az=300)
        guess = syn.pose7_from_xyyaw(g)
        pr = r.register_frame_core(scan, guess, 3.0, 1.0 / 3.0, 0.4)
        po, it = o.register_frame_core(scan, guess, 3.0, 1.0 / 3.0, 0.4)
        dt, da = pose_delta(pr, po)
        assert 1 < it < 500
        assert dt < 1e-9 and da < 1e-10, (seed, dt, da)
    # empty map: the initial guess comes back untouched
    e = ref.RefMap(0.8, 100.0, 20, 20, BASIC_LABELS)
    assert np.allclose(e.register_frame_core(scan, guess, 3.0, 1.0 / 3.0, 0.4), guess, atol=1e-15)


@pytest.mark.parametrize("variant", ["odometry", "360", "raw", "gt"])
def test_pipeline_sequence(ref, orc, variant):
    """sageICP::RegisterFrame over a drive with the parameters of each launch file: Preprocess, Voxelize, GetAdaptiveThreshold
    (stateful), GetPredictionModel, the ICP, UpdateModelDeviation, Update, poses — frame by frame: identical query clouds, poses
    to rounding, identical adaptive-threshold state, and at the end the identical local map IN ORDER."""
    from sage_icp_b200 import synthetic as syn
    from sage_icp_b200.config import launch_config
    over = {"odometry": {}, "360": dict(voxel_size_map=1.0, sem_th=0.8), "raw": dict(sem_th=0.2, local_map_range=40.0),
            "gt": dict(sem_th=0.05)}[variant]
    cfg = launch_config(**over)
    rp, op = ref.RefPipeline(cfg), orc.OraclePipeline(cfg, evict_faithful=True)
    n = 30
    traj = syn.trajectory(n)
    for i in range(n):
        scan = syn.make_scan(100 + i, tuple(traj[i]), n_beams=32, n_az=900)
        pr = rp.register_frame(scan)
        po, _, _ = op.register_frame(scan)
        dt, da = pose_delta(pr, po)
        assert dt < 1e-8 and da < 1e-9, (i, dt, da)
        assert np.array_equal(rp.last_source(), op.last_source()), i
        assert rp.has_moved() == op.has_moved()
    assert np.abs(rp.poses() - op.poses()).max() < 1e-8
    assert np.allclose(rp.prediction_model(), op.prediction_model(), atol=1e-9)
    a, b = rp.local_map(), op.local_map()
    assert a.shape == b.shape and np.array_equal(a[:, 3], b[:, 3]) and np.allclose(a, b, atol=1e-7, rtol=0)
    # GetAdaptiveThreshold is stateful (each call may add a sample): call it once on both
    assert rp.adaptive_threshold() == pytest.approx(op.adaptive_threshold(), rel=1e-7)
    s_r, d_r = rp.voxelize(scan)
    s_o, d_o = op.voxelize(scan)
    assert np.array_equal(s_r, s_o) and np.array_equal(d_r, d_o)
    rp.reset(); op.reset()
    assert len(rp.poses()) == 0 and len(rp.local_map()) == 0
    assert np.allclose(rp.register_frame(scan), op.register_frame(scan)[0])


def test_pipeline_with_deskew(ref, orc):
    """RegisterFrame(frame, timestamps) with deskew = true: no de-skew until three poses exist, then DeSkewScan with the last two."""
    from sage_icp_b200 import synthetic as syn
    from sage_icp_b200.config import launch_config
    cfg = launch_config(deskew=True)
    rp, op = ref.RefPipeline(cfg), orc.OraclePipeline(cfg, evict_faithful=True)
    traj = syn.trajectory(10)
    for i in range(10):
        scan = syn.make_scan(2000 + i, tuple(traj[i]), n_beams=32, n_az=600)
        ts = (np.arange(len(scan)) % 600) / 600.0
        pr = rp.register_frame(scan, ts)
        po, _, _ = op.register_frame(scan, ts)
        dt, da = pose_delta(pr, po)
        assert dt < 1e-8 and da < 1e-9, (i, dt, da)
        assert np.allclose(rp.last_source(), op.last_source(), atol=1e-9), i
    start, finish = op.poses()[-2], op.poses()[-1]
    assert np.allclose(ref.deskew(scan, ts, start, finish), orc.deskew(scan, ts, start, finish), atol=1e-12)


@pytest.mark.parametrize("dy_th,seed", [(0.5, 3), (0.05, 4), (3.0, 5)])
def test_dynamic_vehicle_filter_keeps_the_same_points(ref, orc, dy_th, seed):
    """Preprocess with dynamic_vehicle_filter = true.  The reference's own filter logic (classification, per-cluster landmark
    count against int(dy_th * cluster size), early exit) runs over a stand-in for PCL's Euclidean clustering and FLANN radius
    search, so this pins the SET of kept points: non-vehicle points first, in input order, identical; the re-admitted vehicle
    points equal as a set (the reference emits them cluster by cluster, the oracle and the GPU path in input order —
    DESIGN.md section 6)."""
    from sage_icp_b200.config import launch_config
    cfg = launch_config(dynamic_vehicle_filter=True, dynamic_vehicle_filter_th=dy_th)
    scan = _scan(seed, beams=48, az=1000)
    a, b = ref.preprocess(cfg, scan), orc.preprocess_dynamic(cfg, scan)
    assert a.shape == b.shape
    veh = [10, 11, 13, 15, 16, 18, 20]
    n_plain = (~np.isin(b[:, 3], veh)).sum()
    assert np.array_equal(a[:n_plain], b[:n_plain])
    assert not np.isin(a[:n_plain, 3], veh).any() and np.isin(a[n_plain:, 3], veh).all()
    key = lambda x: x[np.lexsort(x.T[::-1])]
    assert np.array_equal(key(a[n_plain:]), key(b[n_plain:]))
    if dy_th == 0.5:
        assert 0 < len(a) - n_plain
    # with the oracle's cluster_order switch (clusters by descending size, input order inside a cluster) the ORDER matches as well
    assert np.array_equal(a, orc.preprocess_dynamic(cfg, scan, cluster_order=True))


def test_pipeline_with_the_dynamic_vehicle_filter(ref, orc):
    """sageICP::RegisterFrame with dynamic_vehicle_filter = true (three of the four launch files): with the oracle emitting the
    re-admitted vehicle points in the reference's cluster order, whole drives agree — query clouds identical, poses to rounding."""
    from sage_icp_b200 import synthetic as syn
    from sage_icp_b200.config import launch_config
    cfg = launch_config(dynamic_vehicle_filter=True)
    rp, op = ref.RefPipeline(cfg), orc.OraclePipeline(cfg, evict_faithful=True)
    op.set_dynamic_cluster_order(True)
    traj = syn.trajectory(12)
    for i in range(12):
        scan = syn.make_scan(700 + i, tuple(traj[i]), n_beams=32, n_az=800)
        pr = rp.register_frame(scan)
        po, _, _ = op.register_frame(scan)
        dt, da = pose_delta(pr, po)
        assert dt < 1e-8 and da < 1e-9, (i, dt, da)
        assert np.array_equal(rp.last_source(), op.last_source()), i
    a, b = rp.local_map(), op.local_map()
    assert a.shape == b.shape and np.array_equal(a[:, 3], b[:, 3]) and np.allclose(a, b, atol=1e-7, rtol=0)


def test_transform_to_last_frame(ref, orc, cfg):
    rng = np.random.default_rng(2)
    pts = np.c_[rng.uniform(-20, 20, (500, 3)), rng.integers(0, 100, 500).astype(float)]
    last, cur = orc.se3_exp([1, 2, 0.1, 0.01, -0.02, 0.3]), orc.se3_exp([2, 2.1, 0.1, 0.0, -0.01, 0.35])
    rp = ref.RefPipeline(cfg)
    out = rp.transform_to_last_frame(last, cur, pts)
    T = orc.se3_mul(orc.se3_inverse(last), cur)
    exp = np.array([orc.se3_act(T, p[:3]) for p in pts])
    assert np.allclose(out[:, :3], exp, atol=1e-12) and np.array_equal(out[:, 3], pts[:, 3])


def test_committed_golden_vectors_are_what_the_reference_build_computes(ref, cfg):
    """tests/golden/*.npz were generated from the oracle; the GPU tests check the CUDA path against them on the GPU box.  Here
    the reference's own code reproduces them: front end bit for bit, map contents bit for bit, matched targets bit for bit,
    poses to rounding — so the fixtures are reference outputs in all but provenance."""
    import os
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    scan_of = lambda xyz, label: np.c_[xyz.astype(np.float64), label.astype(np.float64)]
    g = np.load(os.path.join(G, "frontend.npz"))
    cropped = ref.preprocess(cfg, scan_of(g["xyz"], g["label"]))
    assert np.array_equal(cropped, g["cropped"])
    ds = ref.voxel_downsample(cfg, cropped, 0.5)
    assert np.array_equal(ds, g["downsample"]) and np.array_equal(ref.voxel_downsample(cfg, ds, 1.5), g["source"])

    g = np.load(os.path.join(G, "core.npz"))
    m = ref.RefMap(0.8, 100.0, 20, 20, BASIC_LABELS)
    m.add_points(scan_of(g["map_xyz"], g["map_label"]))
    keys, counts, vox = m.dump()
    order = np.lexsort(keys.T[::-1])
    assert np.array_equal(keys[order], g["keys"]) and np.array_equal(counts[order], g["counts"])
    assert np.array_equal(vox[order], g["voxels"].astype(np.float64))
    src, tgt = m.get_correspondences(g["queries"], 1.5, 0.4)
    assert np.array_equal(src, g["queries"][g["matched_idx"]]) and np.array_equal(tgt, g["targets"])
    pose = m.register_frame_core(g["queries"], g["guess"], 3.0, 1.0 / 3.0, 0.4)
    assert np.allclose(pose, g["pose"], atol=1e-10)

    g = np.load(os.path.join(G, "sequence.npz"))
    p = ref.RefPipeline(cfg)
    for i in range(len(g["xyz"])):
        pose = p.register_frame(scan_of(g["xyz"][i], g["label"][i]))
        assert np.allclose(pose, g["poses"][i], atol=1e-9), i
        assert len(p.last_source()) == g["n_source"][i]
    assert len(p.local_map()) == int(g["map_points"])


def test_long_drive_with_eviction(ref, orc):
    """90 full-size scans (64 x 1875 rays, which the ICP tracks), 1 m per frame, 40 m map horizon: the map turns over, so the erase-while-iterating sweep, the
    re-use of the robin_map's buckets and the adaptive threshold's accumulated state all matter.  Poses stay equal to rounding
    for the whole drive and the final local map is the same, in order."""
    from sage_icp_b200 import synthetic as syn
    from sage_icp_b200.config import launch_config
    cfg = launch_config(local_map_range=40.0)
    rp, op = ref.RefPipeline(cfg), orc.OraclePipeline(cfg, threads=orc.max_threads(), evict_faithful=True)
    n = 90
    traj = syn.trajectory(n)
    worst = 0.0
    for i in range(n):
        scan = syn.make_scan(1000 + i, tuple(traj[i]), n_beams=64, n_az=1875)
        pr = rp.register_frame(scan)
        po, _, _ = op.register_frame(scan)
        dt, da = pose_delta(pr, po)
        worst = max(worst, dt)
        assert dt < 1e-7 and da < 1e-8, (i, dt, da)
        if i % 15 == 0:
            assert np.array_equal(rp.last_source(), op.last_source()), i
    a, b = rp.local_map(), op.local_map()
    assert a.shape == b.shape
    far = np.linalg.norm(a[:, :3] - rp.poses()[-1][:3], axis=1) > 40.0 + 2.0
    assert rp.poses()[-1][0] - rp.poses()[0][0] > 60.0 and far.mean() < 0.2  # the vehicle out-drove the horizon: most old voxels are gone
    assert np.array_equal(a[:, 3], b[:, 3]) and np.allclose(a, b, atol=1e-6, rtol=0)
    assert rp.adaptive_threshold() == pytest.approx(op.adaptive_threshold(), rel=1e-6)
