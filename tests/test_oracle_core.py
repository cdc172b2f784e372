"""Oracle pinning, part 3: the restated core algorithms (oracle/sage_oracle.hpp) against independent numpy restatements
of the reference source (cited per test), plus the domain's own properties."""
import numpy as np
import pytest

BASIC_LABELS = [40, 44, 48, 49, 50, 70, 72]
IDENT = np.array([0, 0, 0, 0, 0, 0, 1.0])


def _np_keys(p, vs):
    return np.trunc(p[:, :3] / vs).astype(np.int64)  # static_cast<int>: truncation toward zero


def test_preprocess_range_branch(orc):
    """core/Preprocessing.cpp:173-187: strict both sides, far labels zeroed, order kept."""
    rng = np.random.default_rng(0)
    pts = np.c_[rng.normal(0, 40, (5000, 3)), rng.choice([40.0, 50.0, 81.0], 5000)]
    pts[0, :3] = [5.0, 0, 0]      # norm == min_range: dropped (strict >)
    pts[1, :3] = [100.0, 0, 0]    # norm == max_range: dropped (strict <)
    pts[2, :3] = [50.0, 0, 0]     # norm == label_max_range: label kept (strict >)
    out = orc.preprocess(pts, 100.0, 5.0, 50.0)
    nrm = np.sqrt((pts[:, 0] ** 2 + pts[:, 1] ** 2) + pts[:, 2] ** 2)
    keep = (nrm < 100.0) & (nrm > 5.0)
    exp = pts[keep].copy()
    exp[nrm[keep] > 50.0, 3] = 0.0
    assert np.array_equal(out, exp)
    assert not keep[0] and not keep[1] and keep[2]
    assert out[0, 3] == pts[2, 3]  # point 2 is the first survivor; its label is kept


def _py_add_point(block, p, basic, critical):
    """VoxelBlock::AddPoint — core/VoxelHashMap.hpp:45-70, transcribed rule by rule."""
    if len(block) < basic:
        block.append(p)
        return
    label = int(p[3])
    if label == 0:
        return
    if label in BASIC_LABELS:
        for i, q in enumerate(block):
            if int(q[3]) == 0:
                block[i] = p
                return
        return
    if len(block) < basic + critical:
        block.append(p)
        return
    for i, q in enumerate(block):
        if int(q[3]) == 0:
            block[i] = p
            return


@pytest.mark.parametrize("basic,critical", [(20, 20), (3, 2), (1, 0), (4, 0)])
def test_add_points_rule_table(orc, basic, critical):
    rng = np.random.default_rng(1)
    pts = np.c_[rng.uniform(-1.9, 1.9, (4000, 3)), rng.choice([0, 0, 40, 50, 70, 80, 81, 10, 252], 4000).astype(float)]
    m = orc.OracleMap(0.8, 100.0, basic, critical, BASIC_LABELS)
    m.add_points(pts)
    model = {}
    for p in pts:
        k = tuple(np.trunc(p[:3] / 0.8).astype(int))
        if k in model:
            _py_add_point(model[k], p, basic, critical)
        else:
            model[k] = [p]  # a new voxel takes its first point whatever the label (core/VoxelHashMap.cpp:171)
    keys, counts, vox = m.dump()
    assert len(keys) == len(model)
    for k, c, v in zip(keys, counts, vox):
        exp = np.array(model[tuple(int(x) for x in k)])
        assert c == len(exp) and np.array_equal(v[:c], exp)
    assert (0, 0, 0) in model and any(x < 0 for k in model for x in k)  # voxel 0 is double width: -0.3 -> 0


def test_truncation_toward_zero_keys(orc):
    m = orc.OracleMap(0.8, 100.0, 20, 20, BASIC_LABELS)
    pts = np.array([[-0.3, 0.3, -0.79, 40], [0.79, -0.79, 0.0, 40], [-0.81, 0.81, 1.61, 40], [-1.61, 0.0, 0.0, 40.0]])
    m.add_points(pts)
    keys, counts, _ = m.dump()
    got = {tuple(int(x) for x in k): int(c) for k, c in zip(keys, counts)}
    assert got == {(0, 0, 0): 2, (-1, 1, 2): 1, (-2, 0, 0): 1}


@pytest.mark.parametrize("th", [0.4, 0.05, 1.0])
def test_get_correspondences_brute_force(orc, th):
    """core/VoxelHashMap.cpp:48-130 against a numpy brute force over the 27-voxel neighbourhood, including the strict-'<'
    first-wins tie-break in enumeration order (x outer, z inner, stored order) and the unweighted acceptance test."""
    rng = np.random.default_rng(2)
    vs = 0.8
    pts = np.c_[rng.uniform(-4, 4, (3000, 3)), rng.choice([0, 40, 50, 81], 3000).astype(float)]
    m = orc.OracleMap(vs, 100.0, 20, 20, BASIC_LABELS)
    m.add_points(pts)
    keys, counts, vox = m.dump()
    stored = {tuple(int(x) for x in k): v[:c] for k, c, v in zip(keys, counts, vox)}
    q = np.c_[rng.uniform(-5, 5, (400, 3)), rng.choice([0, 40, 50, 81, 10], 400).astype(float)]
    max_dist = 0.6
    src, tgt, qidx = m.get_correspondences(q, max_dist, th)
    got = {int(i): t for i, t in zip(qidx, tgt)}
    for i, p in enumerate(q):
        k = np.trunc(p[:3] / vs).astype(int)
        best, best_pt = np.inf, None
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    for n in stored.get((k[0] + dx, k[1] + dy, k[2] + dz), []):
                        d = ((n[:3] - p[:3]) ** 2).sum()
                        if int(n[3]) == int(p[3]) or int(n[3] * p[3]) == 0:
                            d *= th
                        if d < best:
                            best, best_pt = d, n
        ok = best_pt is not None and np.linalg.norm(best_pt[:3] - p[:3]) < max_dist
        assert ok == (i in got), i
        if ok:
            d_got = ((got[i][:3] - p[:3]) ** 2).sum() * (th if (int(got[i][3]) == int(p[3]) or int(got[i][3] * p[3]) == 0) else 1.0)
            assert np.isclose(d_got, best, rtol=1e-12, atol=0)
    assert 50 < len(got) < 400


def test_normal_equations_match_numpy(orc):
    """core/Registration.cpp:59-94: J = [I | -hat(s)], r = s - t, w = th^2 / (th + |r|^2)^2."""
    rng = np.random.default_rng(3)
    s = np.c_[rng.normal(0, 20, (500, 3)), np.zeros(500)]
    t = s + np.c_[rng.normal(0, 0.2, (500, 3)), np.zeros(500)]
    th = 0.6
    JTJ, JTr, x, est = orc.align_clouds(s, t, th)
    A, b = np.zeros((6, 6)), np.zeros(6)
    for p, n in zip(s, t):
        r = p[:3] - n[:3]
        hat = np.array([[0, -p[2], p[1]], [p[2], 0, -p[0]], [-p[1], p[0], 0]])
        J = np.c_[np.eye(3), -hat]
        w = th ** 2 / (th + r @ r) ** 2
        A += J.T @ (w * J)
        b += J.T @ (w * r)
    assert np.allclose(JTJ, A, rtol=1e-12, atol=1e-9)
    assert np.allclose(JTr, b, rtol=1e-11, atol=1e-9)
    assert np.allclose(x, np.linalg.solve(A, -b), rtol=1e-8)
    assert np.allclose(est, orc.se3_exp(x), atol=1e-15)


def test_icp_recovers_a_known_transform(orc):
    """Property: registering a copy of the map moved by T^-1 returns T (noise-free, converges below the 1e-4 threshold)."""
    rng = np.random.default_rng(4)
    world = np.c_[rng.uniform(-15, 15, (20000, 2)), rng.uniform(0, 3, 20000) * (rng.random(20000) < 0.3), rng.choice([40.0, 50.0], 20000)]
    world[:, :2] = np.round(world[:, :2], 3)
    m = orc.OracleMap(0.8, 100.0, 20, 20, BASIC_LABELS)
    m.add_points(world)
    T = orc.se3_exp([0.12, -0.08, 0.03, 0.004, -0.003, 0.01])
    Tinv = orc.se3_inverse(T)
    frame = world[::7].copy()
    frame[:, :3] = np.array([orc.se3_act(Tinv, p[:3]) for p in frame])
    pose, it = m.register_frame_core(frame, IDENT, 1.0, 0.3, 0.4)
    assert 1 < it < 100
    assert np.linalg.norm(pose[:3] - T[:3]) < 2e-3 and min(np.linalg.norm(pose[3:] - T[3:]), np.linalg.norm(pose[3:] + T[3:])) < 1e-3
    # empty map -> initial guess, zero iterations (core/Registration.cpp:119)
    e = orc.OracleMap(0.8, 100.0, 20, 20, BASIC_LABELS)
    g = orc.se3_exp([1, 2, 3, 0.1, 0.2, 0.3])
    pose, it = e.register_frame_core(frame, g, 1.0, 0.3, 0.4)
    assert it == 0 and np.array_equal(pose, g)


def test_eviction_faithful_vs_clean(orc):
    """core/VoxelHashMap.cpp:176-184: only points.front() is tested; the erase-while-iterating skip can
    only let far voxels survive, never remove near ones."""
    rng = np.random.default_rng(5)
    pts = np.c_[rng.uniform(-60, 60, (30000, 3)) * [1, 1, 0.05], rng.choice([40.0, 0.0], 30000)]
    a = orc.OracleMap(0.8, 30.0, 20, 20, BASIC_LABELS, evict_faithful=True)
    b = orc.OracleMap(0.8, 30.0, 20, 20, BASIC_LABELS, evict_faithful=False)
    a.add_points(pts); b.add_points(pts)
    origin = np.array([5.0, -3.0, 0.0])
    a.remove_far(origin); b.remove_far(origin)
    ka, ca, va = a.dump(); kb, cb, vb = b.dump()
    sa, sb = {tuple(k) for k in ka.tolist()}, {tuple(k) for k in kb.tolist()}
    assert sb <= sa and len(sb) > 100
    for k, c, v in zip(kb, cb, vb):
        assert ((v[0, :3] - origin) ** 2).sum() <= 30.0 ** 2
    extra = sa - sb
    front = {tuple(k): v[0, :3] for k, v in zip(ka.tolist(), va)}
    assert all(((front[k] - origin) ** 2).sum() > 30.0 ** 2 for k in extra)
    a.remove_far(origin)  # a second sweep catches (most of) the skipped ones
    assert a.num_voxels() <= len(sa)


def test_adaptive_threshold_and_first_frames(orc, cfg):
    """pipeline/sageICP.cpp:54-121, core/Threshold.cpp:29-50, restated in numpy over a short drive."""
    from sage_icp_b200 import synthetic as syn
    p = orc.OraclePipeline(cfg, evict_faithful=False)
    traj = syn.trajectory(8)
    sse, n, dev = 0.0, 0, IDENT.copy()
    poses = []
    for i in range(8):
        # expected sigma BEFORE the frame is registered
        moved = len(poses) > 0 and np.linalg.norm(orc.se3_mul(orc.se3_inverse(poses[0]), poses[-1])[:3]) > 5 * cfg.min_motion_th
        if moved:
            theta = orc.rotation_angle(dev)
            err = np.linalg.norm(dev[:3]) + 2.0 * cfg.max_range * np.sin(theta / 2.0)
            if err > cfg.min_motion_th:
                sse += err * err; n += 1
            sigma = cfg.initial_threshold if n < 1 else np.sqrt(sse / n)
        else:
            sigma = cfg.initial_threshold
        pred = IDENT if len(poses) < 2 else orc.se3_mul(orc.se3_inverse(poses[-2]), poses[-1])
        assert np.allclose(p.prediction_model(), pred, atol=1e-15)
        guess = orc.se3_mul(poses[-1] if poses else IDENT, pred)
        scan = syn.make_scan(200 + i, tuple(traj[i]), n_beams=32, n_az=900)
        pose, t_icp, t_all = p.register_frame(scan)
        assert p.last_sigma() == pytest.approx(sigma, rel=1e-12)
        assert 0 <= t_icp <= t_all
        if i == 0:
            assert np.array_equal(pose, IDENT) and p.last_iterations() == 0  # empty map: pose = guess = identity
        dev = orc.se3_mul(orc.se3_inverse(guess), pose)
        poses.append(pose)
    assert np.allclose(p.poses(), np.array(poses))
    assert poses[-1][0] > 0.7 * (traj[-1][0] - traj[0][0]) and abs(poses[-1][1]) < 0.2  # it tracks the drive (the first frames lag: no motion prior yet)
    p.reset()
    assert len(p.poses()) == 0 and p.map().num_voxels() == 0


def test_deskew_matches_numpy(orc):
    """core/Deskew.cpp:36-50: p_i <- exp((t_i - 0.5) * log(start^-1 * finish)) * p_i."""
    rng = np.random.default_rng(6)
    frame = np.c_[rng.normal(0, 20, (300, 3)), rng.choice([40.0, 0.0], 300)]
    ts = rng.random(300)
    a, b = orc.se3_exp([1, 0.2, 0, 0, 0, 0.05]), orc.se3_exp([2.1, 0.3, 0.01, 0.001, 0, 0.09])
    out = orc.deskew(frame, ts, a, b)
    delta = orc.se3_log(orc.se3_mul(orc.se3_inverse(a), b))
    for i in range(300):
        assert np.allclose(out[i, :3], orc.se3_act(orc.se3_exp((ts[i] - 0.5) * delta), frame[i, :3]), atol=1e-13)
    assert np.array_equal(out[:, 3], frame[:, 3])


def test_key_frame_grid_and_overlap_against_numpy(orc):
    """utils::EigenToGridMap / compute_occ_overlap (ros/ros2/Utils.hpp:220-258) restated in the oracle vs a numpy restatement:
    inclusive bounds, the upper bound added before the division, truncation toward zero, cells set to 1; overlap = |s & t| / |s|."""
    rng = np.random.default_rng(3)
    bounds = [[-51.2, 51.2], [-51.2, 51.2], [-4.0, 2.4]]  # ros/launch/odometry.launch.py:88
    H, W = 128, 128                                        # :90
    pts = np.c_[rng.uniform(-60, 60, (20000, 2)), rng.uniform(-6, 4, 20000), rng.integers(0, 100, 20000)].astype(float)
    pts[:50, 0] = 51.2    # on the inclusive upper bound: index W -> rejected by the range check
    pts[50:100, 1] = -51.2
    g = orc.grid_map(pts, bounds, H, W)
    ref = np.zeros((H, W), np.int32)
    xr, yr = (bounds[0][1] - bounds[0][0]) / W, (bounds[1][1] - bounds[1][0]) / H
    for x, y, z, _ in pts:
        if x < bounds[0][0] or x > bounds[0][1] or y < bounds[1][0] or y > bounds[1][1] or z < bounds[2][0] or z > bounds[2][1]:
            continue
        ox, oy = int((x + bounds[0][1]) / xr), int((y + bounds[1][1]) / yr)
        if 0 <= ox < W and 0 <= oy < H:
            ref[oy, ox] = 1
    assert np.array_equal(g, ref) and 0 < g.sum() < H * W
    g2 = orc.grid_map(pts[::2] + [0.7, -0.3, 0, 0], bounds, H, W)
    assert orc.occ_overlap(g, g2) == pytest.approx((g & g2).sum() / g.sum())
    assert orc.occ_overlap(g, g) == 1.0
    assert np.isnan(orc.occ_overlap(np.zeros_like(g), g))  # 0 / 0, as the reference
