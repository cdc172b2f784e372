"""Oracle pinning, part 1: the restated Sophus/Eigen arithmetic (oracle/se3.hpp) against independent implementations —
scipy.linalg.expm/logm on 4x4 matrices, scipy Rotation, numpy solve.  The reference pulls this arithmetic from
un-vendored third-party code, so these checks are what anchors the oracle's SE(3)."""
import numpy as np
import pytest
from scipy.linalg import expm, logm
from scipy.spatial.transform import Rotation as R


def _hat6(xi):
    u, w = xi[:3], xi[3:]
    M = np.zeros((4, 4))
    M[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]
    M[:3, 3] = u
    return M


def _mat(pose):
    T = np.eye(4)
    T[:3, :3] = R.from_quat(pose[3:]).as_matrix()  # scipy: x, y, z, w — same as the wire format
    T[:3, 3] = pose[:3]
    return T


@pytest.mark.parametrize("scale", [1e-12, 1e-6, 1e-2, 0.5, 2.5])
def test_exp_matches_matrix_exponential(orc, scale):
    rng = np.random.default_rng(int(scale * 1e6) % 1000)
    for _ in range(20):
        xi = rng.normal(size=6) * scale
        xi[:3] *= 10
        T = _mat(orc.se3_exp(xi))
        assert np.allclose(T, expm(_hat6(xi)), atol=1e-12, rtol=1e-12)


def test_log_inverts_exp_and_matches_logm(orc):
    rng = np.random.default_rng(1)
    for _ in range(50):
        xi = rng.normal(size=6) * np.array([3, 3, 3, 0.8, 0.8, 0.8])
        pose = orc.se3_exp(xi)
        assert np.allclose(orc.se3_log(pose), xi, atol=1e-11)
        L = np.real(logm(_mat(pose)))
        assert np.allclose(L, _hat6(orc.se3_log(pose)), atol=1e-9)
    assert np.allclose(orc.se3_log(orc.se3_exp(np.zeros(6))), 0)
    tiny = np.array([1e-5, -2e-5, 3e-5, 1e-13, -1e-13, 2e-13])
    assert np.allclose(orc.se3_log(orc.se3_exp(tiny)), tiny, atol=1e-18, rtol=1e-9)


def test_group_operations(orc):
    rng = np.random.default_rng(2)
    for _ in range(30):
        a = orc.se3_exp(rng.normal(size=6))
        b = orc.se3_exp(rng.normal(size=6))
        p = rng.normal(size=3) * 20
        assert np.allclose(_mat(orc.se3_mul(a, b)), _mat(a) @ _mat(b), atol=1e-12)
        assert np.allclose(_mat(orc.se3_inverse(a)), np.linalg.inv(_mat(a)), atol=1e-12)
        assert np.allclose(orc.se3_act(a, p), (_mat(a) @ np.r_[p, 1])[:3], atol=1e-12)
        assert abs(np.linalg.norm(orc.se3_mul(a, b)[3:]) - 1) < 1e-15  # renormalised like Sophus


def test_rotation_angle_is_angle_axis_angle(orc):
    rng = np.random.default_rng(3)
    for _ in range(30):
        w = rng.normal(size=3)
        w *= rng.uniform(0, np.pi - 1e-3) / np.linalg.norm(w)
        pose = orc.se3_exp(np.r_[0, 0, 0, w])
        assert orc.rotation_angle(pose) == pytest.approx(np.linalg.norm(w), abs=1e-12)
    assert orc.rotation_angle(np.array([0, 0, 0, 0, 0, 0, 1.0])) == 0.0
    assert orc.rotation_angle(np.array([0, 0, 0, 0, 0, 0, -1.0])) == 0.0  # q and -q: same rotation, angle in [0, pi]


def test_ldlt_solve_matches_numpy(orc):
    rng = np.random.default_rng(4)
    for cond in (1.0, 1e3, 1e6):
        for _ in range(10):
            J = rng.normal(size=(40, 6)) * np.logspace(0, np.log10(cond) / 2, 6)
            A = J.T @ J
            b = rng.normal(size=6)
            x = orc.ldlt6_solve(A, b)
            assert np.allclose(A @ x, b, rtol=1e-9, atol=1e-9 * np.abs(b).max() * cond)
            assert np.allclose(x, np.linalg.solve(A, b), rtol=1e-6 * cond ** 0.5)
