"""The search kernel ranks candidates on 16-byte f32 records, prunes neighbour voxels by a bounding-box bound and falls
back to the reference's f64 scan when the f32 ranking is ambiguous.  These cases attack exactly those three mechanisms:
the result must stay bit-identical to the oracle's plain f64 scan of all 27 voxels (core/VoxelHashMap.cpp:48-130)."""
import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("search_mode")]

BASIC_LABELS = [40, 44, 48, 49, 50, 70, 72]


def _pair(orc, voxel_size=0.8, basic=20, critical=20):
    import sage_icp_b200 as sg
    g = sg.SageMap(voxel_size, 1e9, basic, critical, BASIC_LABELS)
    o = orc.OracleMap(voxel_size, 1e9, basic, critical, BASIC_LABELS, evict_faithful=False)
    return g, o


def _check(g, o, q, max_dist, th):
    tgt, matched = g.get_correspondences(q, max_dist, th)
    _, tgt_o, qidx = o.get_correspondences(q, max_dist, th)
    m_o = np.zeros(len(q), bool)
    m_o[qidx] = True
    assert np.array_equal(matched, m_o)
    assert np.array_equal(tgt[matched], tgt_o)
    return int(matched.sum())


@pytest.mark.parametrize("th", [0.4, 1.0, 2.5, 0.05])
def test_exact_ties_on_a_lattice(orc, th):
    """Map points on a regular lattice, queries on lattice midpoints and on lattice points: masses of exact f64 ties that
    only the first-wins enumeration order resolves (core/VoxelHashMap.cpp:57-63,89)."""
    g, o = _pair(orc)
    ax = np.arange(-16, 16) * 0.25  # exactly representable, 3.2 points per voxel edge
    X, Y, Z = np.meshgrid(ax, ax, ax[12:20], indexing="ij")
    rng = np.random.default_rng(0)
    pts = np.stack([X.ravel(), Y.ravel(), Z.ravel(), rng.choice([0, 40, 50, 80], X.size)], 1)
    rng.shuffle(pts)
    g.add_points(pts)
    o.add_points(pts)
    q = pts[rng.choice(len(pts), 4000, replace=False)].copy()
    q[:2000, :3] += 0.125  # equidistant from 8 lattice points
    q[2000:3000, 0] += 0.125  # equidistant from 2
    q[:, 3] = rng.choice([0, 40, 50, 81, 10], len(q))
    n = _check(g, o, q, 1.0, th)
    assert n > 3000
    scanned, probes, exact, heavy = g.search_work(q, 1.0, th)
    assert exact >= 1000  # the tie cases really went down the f64 path


def test_duplicate_points_and_zero_distance(orc):
    g, o = _pair(orc)
    rng = np.random.default_rng(1)
    base = rng.uniform(-3, 3, (300, 3))
    pts = np.concatenate([base, base, base])  # every point three times, different labels
    pts = np.c_[pts, np.r_[np.full(300, 40.0), np.full(300, 0.0), np.full(300, 81.0)]]
    g.add_points(pts)
    o.add_points(pts)
    q = np.c_[base, rng.choice([40, 81, 0, 10], 300)]  # distance exactly 0 to three candidates
    assert _check(g, o, q, 0.5, 0.4) == 300


def test_non_integer_and_negative_labels_take_the_exact_path(orc):
    g, o = _pair(orc)
    rng = np.random.default_rng(2)
    pts = np.c_[rng.uniform(-4, 4, (4000, 3)), rng.choice([0.0, 0.5, 40.0, 40.7, -3.0, 1e9, 0.001], 4000)]
    g.add_points(pts)
    o.add_points(pts)
    q = np.c_[rng.uniform(-4, 4, (3000, 3)), rng.choice([0.0, 0.5, 40.0, 40.7, -3.0, 2.0, 1e-3], 3000)]
    assert _check(g, o, q, 2.0, 0.4) > 2000


@pytest.mark.parametrize("offset", [(0.0, 0.0, 0.0), (5000.3, -7321.9, 12.7), (-0.4, 0.4, -0.4), (1.3e5, 2.0e5, -900.0)])
def test_far_from_origin_and_around_voxel_zero(orc, offset):
    """f32 records are voxel-relative, so parity must not depend on how far from the origin the scene is; offset
    (-0.4, 0.4, -0.4) puts the scene across the double-width voxel 0 (truncation toward zero)."""
    from sage_icp_b200 import synthetic as syn
    g, o = _pair(orc)
    pts = syn.sample_street_map(200_000, 5, -40.0, 40.0)
    pts[:, :3] += np.array(offset)
    o.add_points(pts)
    g.add_points(pts)
    scan = syn.make_scan(7, (0.0, 0.0, 0.0), n_beams=32, n_az=400)
    r = np.linalg.norm(scan[:, :3], axis=1)
    q = scan[(r > 3) & (r < 45)].copy()
    q[:, :3] += np.array(offset) + np.array([0.2, -0.1, syn.SENSOR_HEIGHT])
    assert _check(g, o, q, 1.5, 0.4) > 3000
    scanned, probes, exact, heavy = g.search_work(q, 1.5, 0.4)
    occ, cand = g.nn_stats(q)
    assert scanned < cand  # pruning skipped part of the 27-voxel neighbourhood ...
    assert exact < 0.01 * len(q)  # ... and the f64 fallback stayed rare


@pytest.mark.parametrize("voxel_size,basic,critical", [(0.3, 5, 3), (2.0, 40, 40), (1.0, 1, 0)])
def test_other_voxel_geometries(orc, voxel_size, basic, critical):
    g, o = _pair(orc, voxel_size, basic, critical)
    rng = np.random.default_rng(3)
    pts = np.c_[rng.normal(0, 3, (60_000, 3)), rng.choice([0, 40, 50, 80, 81], 60_000)]
    g.add_points(pts)
    o.add_points(pts)
    q = np.c_[rng.normal(0, 3.5, (8000, 3)), rng.choice([0, 40, 50, 80, 99], 8000)]
    assert _check(g, o, q, 1.0 * voxel_size, 0.4) > 1000


def test_sparse_map_queries_far_from_any_point(orc):
    """Empty home voxels, lone neighbours, empty neighbourhoods (the reference reads an uninitialised vector there;
    oracle and GPU define it as 'no correspondence')."""
    g, o = _pair(orc)
    rng = np.random.default_rng(4)
    pts = np.c_[rng.uniform(-50, 50, (3000, 3)), rng.choice([0, 40, 81], 3000)]
    g.add_points(pts)
    o.add_points(pts)
    q = np.c_[rng.uniform(-55, 55, (20000, 3)), rng.choice([0, 40, 81], 20000)]
    n = _check(g, o, q, 6.0, 0.4)
    assert 0 < n < len(q)


def test_sem_th_zero_and_negative_still_match(orc):
    """Degenerate sem_th disables the f32 ranking altogether (metric no longer orders like its bit pattern)."""
    g, o = _pair(orc)
    rng = np.random.default_rng(6)
    pts = np.c_[rng.uniform(-3, 3, (5000, 3)), rng.choice([0, 40, 81], 5000)]
    g.add_points(pts)
    o.add_points(pts)
    q = np.c_[rng.uniform(-3, 3, (2000, 3)), rng.choice([0, 40, 81, 10], 2000)]
    for th in (0.0, -1.0):
        _check(g, o, q, 2.0, th)
