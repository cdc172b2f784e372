"""Committed golden vectors (tests/golden/*.npz, made by tools/make_golden.py from the oracle): the CPU half pins the oracle
against regressions, the GPU half pins the CUDA path through the C ABI on a box where only these files exist."""
import os

import numpy as np
import pytest

from conftest import POSE_TOL_M, POSE_TOL_RAD, pose_delta

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BASIC_LABELS = [40, 44, 48, 49, 50, 70, 72]


def _load(name):
    return np.load(os.path.join(G, name))


def _scan(xyz, label):
    return np.c_[xyz.astype(np.float64), label.astype(np.float64)]


# ---- CPU: oracle == golden ---------------------------------------------------------------------------------

def test_oracle_se3_golden(orc):
    g = _load("se3.npz")
    for i, xi in enumerate(g["xi"]):
        assert np.allclose(orc.se3_exp(xi), g["exp"][i], atol=1e-14)
        assert np.allclose(orc.se3_log(g["exp"][i]), g["log"][i], atol=1e-13)
        assert np.allclose(orc.se3_mul(g["exp"][i], g["exp"][(i + 1) % 32]), g["mul"][i], atol=1e-14)
        assert np.allclose(orc.se3_inverse(g["exp"][i]), g["inv"][i], atol=1e-14)


def test_oracle_frontend_golden(orc, cfg):
    g = _load("frontend.npz")
    scan = _scan(g["xyz"], g["label"])
    cropped = orc.preprocess(scan, cfg.max_range, cfg.min_range, cfg.label_max_range)
    assert np.array_equal(cropped, g["cropped"])
    ds = orc.voxel_downsample(cfg, cropped, 0.5)
    assert np.array_equal(ds, g["downsample"])
    assert np.array_equal(orc.voxel_downsample(cfg, ds, 1.5), g["source"])
    assert len(g["source"]) < len(g["downsample"]) < len(g["cropped"]) < len(scan)


def _golden_map_dump(g):
    return g["keys"], g["counts"], g["voxels"].astype(np.float64)


def _sorted_dump(dump):
    keys, counts, vox = dump
    order = np.lexsort(keys.T[::-1])
    return keys[order], counts[order], vox[order]


def test_oracle_core_golden(orc):
    g = _load("core.npz")
    m = orc.OracleMap(0.8, 100.0, 20, 20, BASIC_LABELS, evict_faithful=False)
    m.add_points(_scan(g["map_xyz"], g["map_label"]))
    k, c, v = _sorted_dump(m.dump())
    gk, gc, gv = _golden_map_dump(g)
    assert np.array_equal(k, gk) and np.array_equal(c, gc) and np.array_equal(v, gv)
    _, tgt, qidx = m.get_correspondences(g["queries"], 1.5, 0.4)
    assert np.array_equal(qidx, g["matched_idx"]) and np.array_equal(tgt, g["targets"])
    pose, it = m.register_frame_core(g["queries"], g["guess"], 3.0, 1.0 / 3.0, 0.4)
    assert it == int(g["iters"]) and np.allclose(pose, g["pose"], atol=1e-11)


def test_oracle_sequence_golden(orc, cfg):
    g = _load("sequence.npz")
    p = orc.OraclePipeline(cfg, evict_faithful=False)
    for i in range(len(g["xyz"])):
        pose, _, _ = p.register_frame(_scan(g["xyz"][i], g["label"][i]))
        assert np.allclose(pose, g["poses"][i], atol=1e-10), i
        assert p.last_iterations() == g["iterations"][i] and p.last_sigma() == pytest.approx(g["sigma"][i], rel=1e-10)
        assert len(p.last_source()) == g["n_source"][i] and len(p.last_frame_downsample()) == g["n_downsample"][i]
    assert p.map().num_voxels() == int(g["map_voxels"]) and p.map().num_points() == int(g["map_points"])


# ---- GPU: CUDA path == golden ------------------------------------------------------------------------------

@pytest.mark.gpu
def test_gpu_frontend_golden(cfg):
    import sage_icp_b200 as sg
    g = _load("frontend.npz")
    p = sg.SagePipeline(cfg)
    scan = _scan(g["xyz"], g["label"])
    cropped = p.preprocess(scan)
    assert np.array_equal(cropped, g["cropped"])
    ds = p.voxel_downsample(cropped, 0.5)
    assert np.array_equal(ds, g["downsample"])  # same survivors, same (robin_map) order
    assert np.array_equal(p.voxel_downsample(ds, 1.5), g["source"])
    s, d = p.voxelize(cropped)
    assert np.array_equal(s, g["source"]) and np.array_equal(d, g["downsample"])


@pytest.mark.gpu
def test_gpu_core_golden():
    import sage_icp_b200 as sg
    g = _load("core.npz")
    m = sg.SageMap(0.8, 100.0, 20, 20, BASIC_LABELS)
    m.add_points(_scan(g["map_xyz"], g["map_label"]))
    k, c, v = _sorted_dump(m.dump())
    gk, gc, gv = _golden_map_dump(g)
    assert np.array_equal(k, gk) and np.array_equal(c, gc) and np.array_equal(v, gv)
    tgt, matched = m.get_correspondences(g["queries"], 1.5, 0.4)
    assert np.array_equal(np.flatnonzero(matched), g["matched_idx"]) and np.array_equal(tgt[matched], g["targets"])
    JTJ, JTr, n = m.normal_equations(g["queries"], 1.5, 0.5, 0.4)
    assert n == len(g["matched_idx"])
    assert np.allclose(JTJ, g["JTJ"], rtol=1e-11, atol=1e-9 * np.abs(g["JTJ"]).max())
    assert np.allclose(JTr, g["JTr"], rtol=1e-10, atol=1e-9 * np.abs(g["JTr"]).max())
    pose, it = m.register_frame(g["queries"], g["guess"], 3.0, 1.0 / 3.0, 0.4)
    dt, da = pose_delta(pose, g["pose"])
    assert it == int(g["iters"]) and dt <= POSE_TOL_M and da <= POSE_TOL_RAD, (it, dt, da)


@pytest.mark.gpu
def test_gpu_sequence_golden(cfg):
    import sage_icp_b200 as sg
    g = _load("sequence.npz")
    p = sg.SagePipeline(cfg)
    for i in range(len(g["xyz"])):
        pose, t_icp, t_all = p.register_frame(_scan(g["xyz"][i], g["label"][i]))
        dt, da = pose_delta(pose, g["poses"][i])
        assert dt <= POSE_TOL_M and da <= POSE_TOL_RAD, (i, dt, da)
        assert p.last_iterations() == g["iterations"][i] and p.last_sigma() == pytest.approx(g["sigma"][i], rel=1e-9)
        assert len(p.last_source()) == g["n_source"][i] and len(p.last_frame_downsample()) == g["n_downsample"][i]
    assert p.map().num_voxels() == int(g["map_voxels"]) and p.map().num_points() == int(g["map_points"])
