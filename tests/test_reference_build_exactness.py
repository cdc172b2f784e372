"""The adversarial nearest-neighbour cases of tests/test_gpu_search_exactness.py (exact ties on a lattice, duplicates, fractional /
negative / huge labels, scenes far from the origin and across voxel 0, other voxel geometries, sparse maps with empty
neighbourhoods, degenerate sem_th), here between the oracle and the reference's own GetCorrespondences (oracle/_ref, see
tests/test_reference_build.py): the same pairs in the same order, bit for bit.  With the GPU file this closes the chain
CUDA path == oracle == reference code on the inputs most likely to tell them apart."""
import numpy as np
import pytest

BASIC_LABELS = [40, 44, 48, 49, 50, 70, 72]


@pytest.fixture(scope="module")
def ref():
    from oracle import ref_py
    if not ref_py.available():
        pytest.skip("neither oracle/_ref/libsage_ref.so nor /root/reference is present")
    ref_py.lib()
    return ref_py


def _pair(ref, orc, voxel_size=0.8, basic=20, critical=20):
    return (ref.RefMap(voxel_size, 1e9, basic, critical, BASIC_LABELS),
            orc.OracleMap(voxel_size, 1e9, basic, critical, BASIC_LABELS, evict_faithful=True))


def _check(r, o, q, max_dist, th):
    rs, rt = r.get_correspondences(q, max_dist, th)
    os_, ot, qidx = o.get_correspondences(q, max_dist, th)
    assert np.array_equal(rs, os_) and np.array_equal(rt, ot)
    assert np.array_equal(rs, q[qidx])
    return len(qidx)


@pytest.mark.parametrize("th", [0.4, 1.0, 2.5, 0.05])
def test_exact_ties_on_a_lattice(ref, orc, th):
    r, o = _pair(ref, orc)
    ax = np.arange(-16, 16) * 0.25
    X, Y, Z = np.meshgrid(ax, ax, ax[12:20], indexing="ij")
    rng = np.random.default_rng(0)
    pts = np.stack([X.ravel(), Y.ravel(), Z.ravel(), rng.choice([0, 40, 50, 80], X.size)], 1)
    rng.shuffle(pts)
    r.add_points(pts); o.add_points(pts)
    q = pts[rng.choice(len(pts), 4000, replace=False)].copy()
    q[:2000, :3] += 0.125
    q[2000:3000, 0] += 0.125
    q[:, 3] = rng.choice([0, 40, 50, 81, 10], len(q))
    assert _check(r, o, q, 1.0, th) > 3000


def test_duplicate_points_and_zero_distance(ref, orc):
    r, o = _pair(ref, orc)
    rng = np.random.default_rng(1)
    base = rng.uniform(-3, 3, (300, 3))
    pts = np.c_[np.concatenate([base, base, base]), np.r_[np.full(300, 40.0), np.full(300, 0.0), np.full(300, 81.0)]]
    r.add_points(pts); o.add_points(pts)
    q = np.c_[base, rng.choice([40, 81, 0, 10], 300)]
    assert _check(r, o, q, 0.5, 0.4) == 300


def test_non_integer_negative_and_huge_labels(ref, orc):
    r, o = _pair(ref, orc)
    rng = np.random.default_rng(2)
    pts = np.c_[rng.uniform(-4, 4, (4000, 3)), rng.choice([0.0, 0.5, 40.0, 40.7, -3.0, 1e9, 0.001], 4000)]
    r.add_points(pts); o.add_points(pts)
    q = np.c_[rng.uniform(-4, 4, (3000, 3)), rng.choice([0.0, 0.5, 40.0, 40.7, -3.0, 2.0, 1e-3], 3000)]
    assert _check(r, o, q, 2.0, 0.4) > 2000


@pytest.mark.parametrize("offset", [(0.0, 0.0, 0.0), (5000.3, -7321.9, 12.7), (-0.4, 0.4, -0.4), (1.3e5, 2.0e5, -900.0)])
def test_far_from_origin_and_around_voxel_zero(ref, orc, offset):
    from sage_icp_b200 import synthetic as syn
    r, o = _pair(ref, orc)
    pts = syn.sample_street_map(200_000, 5, -40.0, 40.0)
    pts[:, :3] += np.array(offset)
    r.add_points(pts); o.add_points(pts)
    (rk, rc, rp), (ok, oc, op) = r.dump(), o.dump()
    assert np.array_equal(rk, ok) and np.array_equal(rc, oc) and np.array_equal(rp, op)
    scan = syn.make_scan(7, (0.0, 0.0, 0.0), n_beams=32, n_az=400)
    rad = np.linalg.norm(scan[:, :3], axis=1)
    q = scan[(rad > 3) & (rad < 45)].copy()
    q[:, :3] += np.array(offset) + np.array([0.2, -0.1, syn.SENSOR_HEIGHT])
    assert _check(r, o, q, 1.5, 0.4) > 3000


@pytest.mark.parametrize("voxel_size,basic,critical", [(0.3, 5, 3), (2.0, 40, 40), (1.0, 1, 0)])
def test_other_voxel_geometries(ref, orc, voxel_size, basic, critical):
    r, o = _pair(ref, orc, voxel_size, basic, critical)
    rng = np.random.default_rng(3)
    pts = np.c_[rng.normal(0, 3, (60_000, 3)), rng.choice([0, 40, 50, 80, 81], 60_000)]
    r.add_points(pts); o.add_points(pts)
    q = np.c_[rng.normal(0, 3.5, (8000, 3)), rng.choice([0, 40, 50, 80, 99], 8000)]
    assert _check(r, o, q, 1.0 * voxel_size, 0.4) > 1000


def test_sparse_map_queries_far_from_any_point(ref, orc):
    """Empty neighbourhoods: the reference reads a vector it never set (NaN in this build: oracle/shim/Eigen/Core), the oracle
    says 'no correspondence' — the same pairs come out."""
    r, o = _pair(ref, orc)
    rng = np.random.default_rng(4)
    pts = np.c_[rng.uniform(-50, 50, (3000, 3)), rng.choice([0, 40, 81], 3000)]
    r.add_points(pts); o.add_points(pts)
    q = np.c_[rng.uniform(-55, 55, (20000, 3)), rng.choice([0, 40, 81], 20000)]
    assert 0 < _check(r, o, q, 6.0, 0.4) < len(q)


def test_sem_th_zero_and_negative(ref, orc):
    r, o = _pair(ref, orc)
    rng = np.random.default_rng(6)
    pts = np.c_[rng.uniform(-3, 3, (5000, 3)), rng.choice([0, 40, 81], 5000)]
    r.add_points(pts); o.add_points(pts)
    q = np.c_[rng.uniform(-3, 3, (2000, 3)), rng.choice([0, 40, 81, 10], 2000)]
    for th in (0.0, -1.0):
        _check(r, o, q, 2.0, th)
