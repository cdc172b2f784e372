"""The tile search (sage_icp_b200/csrc/search_tile.cuh + tile_sort.cu): queries sorted by 2x2x2-voxel cell, buckets staged through
TMA bulk copies, thread phase + warp phase against shared memory.  Its per-query results must be those of the oracle's f64
27-voxel scan (core/VoxelHashMap.cpp:48-130) whatever the schedule, and a registration must not depend on which kernel ran."""
import numpy as np
import pytest

from conftest import POSE_TOL_M, POSE_TOL_RAD, pose_delta

pytestmark = pytest.mark.gpu

BASIC_LABELS = [40, 44, 48, 49, 50, 70, 72]


def _maps(orc, pts, monkeypatch=None, env=None):
    import sage_icp_b200 as sg
    if env is not None:
        for k in ("SAGE_TILE", "SAGE_TILE_MIN", "SAGE_TILE_FILL", "SAGE_TILE_PERSISTENT", "SAGE_TILE_STAGE", "SAGE_TILE_MINB", "SAGE_TILE_BLOCKS", "SAGE_STEP_EVERYWHERE", "SAGE_TILE_GRAPH"):
            monkeypatch.delenv(k, raising=False)
        if env.get("SAGE_TILE") != "0":
            monkeypatch.setenv("SAGE_TILE_FILL", "0")  # the tile search whatever the density (these tests are about ITS results)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
    g = sg.SageMap(0.8, 1e9, 20, 20, BASIC_LABELS)
    g.add_points(pts)
    return g


def _street(n=400_000):
    from sage_icp_b200 import synthetic as syn
    return syn.sample_street_map(n, 11, -60.0, 60.0)


def _queries(n_beams=64, n_az=500):
    from sage_icp_b200 import synthetic as syn
    scan = syn.make_scan(3, (0.0, 0.0, 0.0), n_beams=n_beams, n_az=n_az)
    return scan, syn.pose7_from_xyyaw((0.25, -0.1, 0.012))


def _check_corr(g, o, q, max_dist, th):
    tgt, matched = g.get_correspondences(q, max_dist, th)
    _, tgt_o, qidx = o.get_correspondences(q, max_dist, th, threads=8)
    m_o = np.zeros(len(q), bool)
    m_o[qidx] = True
    assert np.array_equal(matched, m_o)
    assert np.array_equal(tgt[matched], tgt_o)
    return int(matched.sum())


@pytest.mark.parametrize("env", [{}, {"SAGE_TILE_STAGE": "128"}, {"SAGE_TILE_MINB": "4"}, {"SAGE_TILE_MINB": "8", "SAGE_TILE_BLOCKS": "2"}],
                         ids=["default", "tiny_staging", "128_registers", "64_registers_2_blocks_per_sm"])
def test_tile_correspondences_bit_exact_on_a_full_scan(orc, monkeypatch, env):
    """32 000 queries of a street scan (above the 12 288-query threshold): every query's target equals the oracle's, in the
    caller's order, whether the buckets fit the staging area or are scanned from global memory, for each register budget."""
    pts = _street()
    g = _maps(orc, pts, monkeypatch, env)
    o = orc.OracleMap(0.8, 1e9, 20, 20, BASIC_LABELS, evict_faithful=False)
    o.add_points(pts)
    scan, _ = _queries()
    q = scan.copy()
    q[:, 2] += 1.73
    assert len(q) >= 12288
    assert _check_corr(g, o, q, 3.0, 0.4) > 20000
    scanned, probes, exact, heavy, staged = g.search_work(q, 3.0, 0.4, with_staged=True)
    assert staged > 0  # the buckets really went through the bulk copies
    assert heavy > 0   # ... and some (query, bucket) pairs through the pooled round


def test_tile_registration_equals_legacy_and_is_reproducible(orc, monkeypatch):
    pts = _street()
    scan, guess = _queries()
    o = orc.OracleMap(0.8, 1e9, 20, 20, BASIC_LABELS, evict_faithful=False)
    o.add_points(pts)
    pose_o, it_o = o.register_frame_core(scan, guess, 3.0, 1.0 / 3.0, 0.4, threads=8)
    out = {}
    for name, env in (("tile", {"SAGE_STEP_EVERYWHERE": "0"}), ("tile_every_block_steps", {"SAGE_STEP_EVERYWHERE": "2"}),
                      ("tile_launch_per_iteration", {"SAGE_TILE_PERSISTENT": "0"}), ("legacy", {"SAGE_TILE": "0"})):
        g = _maps(orc, pts, monkeypatch, env)
        p1, it1 = g.register_frame(scan, guess, 3.0, 1.0 / 3.0, 0.4)
        p2, it2 = g.register_frame(scan, guess, 3.0, 1.0 / 3.0, 0.4)
        assert it1 == it2 and np.array_equal(p1, p2), name  # same launch shape -> same bits
        dt, da = pose_delta(p1, pose_o)
        assert it1 == it_o and dt <= POSE_TOL_M and da <= POSE_TOL_RAD, (name, it1, it_o, dt, da)
        out[name] = p1
    assert np.array_equal(out["tile"], out["tile_launch_per_iteration"])  # one cooperative launch == one launch per iteration
    assert np.array_equal(out["tile"], out["tile_every_block_steps"])     # ... == every block taking the step itself
    dt, da = pose_delta(out["tile"], out["legacy"])
    assert dt <= 1e-9 and da <= 1e-10, (dt, da)  # different summation trees only


def test_tile_units_that_do_not_fit_fall_back_to_the_global_search(orc, monkeypatch):
    """Cells 256 cells (409.6 m) apart share a sort key: queries of both land in one unit whose region cannot fit, and the
    unit is searched from global memory.  Scattered queries (one per cell) and a scene across voxel 0 ride along."""
    rng = np.random.default_rng(5)
    a = np.c_[rng.uniform(-6, 6, (60_000, 3)), rng.choice([0, 40, 50, 81], 60_000)]
    b = a.copy()
    b[:, 0] += 409.6
    b[:, 1] -= 2 * 409.6
    pts = np.concatenate([a, b])
    g = _maps(orc, pts, monkeypatch, {"SAGE_TILE_MIN": "1"})
    o = orc.OracleMap(0.8, 1e9, 20, 20, BASIC_LABELS, evict_faithful=False)
    o.add_points(pts)
    qa = np.c_[rng.uniform(-7, 7, (9000, 3)), rng.choice([0, 40, 50, 81, 10], 9000)]
    qb = qa.copy()
    qb[:, 0] += 409.6
    qb[:, 1] -= 2 * 409.6
    q = np.concatenate([qa, qb, np.c_[rng.uniform(-400, 400, (2000, 3)), np.zeros(2000)]])
    q = q[rng.permutation(len(q))]
    assert _check_corr(g, o, q, 2.0, 0.4) > 10000


def test_tile_handles_empty_ragged_and_degenerate_inputs(orc, monkeypatch):
    rng = np.random.default_rng(8)
    pts = np.c_[rng.uniform(-5, 5, (30_000, 3)), rng.choice([0, 40, 81], 30_000)]
    g = _maps(orc, pts, monkeypatch, {"SAGE_TILE_MIN": "1"})
    o = orc.OracleMap(0.8, 1e9, 20, 20, BASIC_LABELS, evict_faithful=False)
    o.add_points(pts)
    for n in (1, 2, 31, 127, 128, 129, 1025):
        q = np.c_[rng.uniform(-6, 6, (n, 3)), rng.choice([0, 40, 81, 0.5], n)]
        _check_corr(g, o, q, 1.5, 0.4)
    q = np.c_[rng.uniform(-6, 6, (500, 3)), rng.choice([0, 40], 500)]
    q[::7, 0] = np.nan
    q[3::11, 1] = np.inf
    q[5::13, 2] = 3e9  # outside the packable key range
    tgt, matched = g.get_correspondences(q, 1.5, 0.4)
    ok = np.isfinite(q[:, :3]).all(1) & (np.abs(q[:, :3]) < 1e6).all(1)
    assert not matched[~ok].any()
    _, tgt_o, qidx = o.get_correspondences(q[ok], 1.5, 0.4)
    m_o = np.zeros(int(ok.sum()), bool)
    m_o[qidx] = True
    assert np.array_equal(matched[ok], m_o) and np.array_equal(tgt[ok][m_o], tgt_o)
    for th in (0.0, -1.0, 2.5):
        _check_corr(g, o, q[ok], 1.5, th)


def test_captured_prep_graph_equals_plain_launches(orc, monkeypatch):
    """The sort + unit-list kernels of a registration are captured into a CUDA graph the second time a scan of the same size arrives
    and replayed from then on (tile_sort.cu).  Different scans and guesses of one size, a different size in between (the graph is
    dropped and rebuilt), host and device-resident entry points: every pose must carry the bits of the plain launches."""
    import torch
    from sage_icp_b200 import synthetic as syn
    pts = _street()
    scans = [syn.make_scan(s, (0.4 * s, 0.0, 0.0), n_beams=64, n_az=500) for s in range(4)]
    guesses = [syn.pose7_from_xyyaw((0.4 * s + 0.2, -0.1 + 0.05 * s, 0.01 * (s - 1))) for s in range(4)]
    small = syn.make_scan(9, (0.0, 0.0, 0.0), n_beams=64, n_az=300)
    order = [0, 1, 2, 3, "small", "small", 3, 2, 1, 0]
    out = {}
    for name, env in (("graph", {}), ("plain", {"SAGE_TILE_GRAPH": "0"})):
        g = _maps(orc, pts, monkeypatch, env)
        poses = []
        for k in order:
            if k == "small":
                poses.append(g.register_frame(small, guesses[0], 3.0, 1.0 / 3.0, 0.4, 8, 0.0))
            elif k % 2:  # scan resident in device memory (the frame pointer changes from call to call)
                d = torch.from_numpy(scans[k]).cuda()
                poses.append(g.register_frame_device(d.data_ptr(), len(scans[k]), guesses[k], 3.0, 1.0 / 3.0, 0.4, 8, 0.0))
                del d
            else:
                poses.append(g.register_frame(scans[k], guesses[k], 3.0, 1.0 / 3.0, 0.4, 8, 0.0))
        out[name] = poses
    for (pg, ig), (pp, ip) in zip(out["graph"], out["plain"]):
        assert ig == ip and np.array_equal(np.asarray(pg), np.asarray(pp))
    o = orc.OracleMap(0.8, 1e9, 20, 20, BASIC_LABELS, evict_faithful=False)
    o.add_points(pts)
    pose_o, it_o = o.register_frame_core(scans[3], guesses[3], 3.0, 1.0 / 3.0, 0.4, threads=8, max_iters=8, est_th=0.0)
    dt, da = pose_delta(np.asarray(out["graph"][3][0]), pose_o)
    assert out["graph"][3][1] == it_o and dt <= POSE_TOL_M and da <= POSE_TOL_RAD
