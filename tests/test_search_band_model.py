"""CPU model of the search kernel's f32 ranking: numpy float32 emulation of the 16-byte search records
(offsets from the voxel origin, labels as exact f32 integers), of the neighbour-relative query, of D32 and the metric, and of the
acceptance band T(gm) = gm + 1.01 (e(gm/s) + e((gm + e(gm/s))/s)) with e(D) = max(1,th)(64u vs sqrt(D) + 12u D) + 1e-11 vs^2.
Claim under test: whenever EXACTLY ONE candidate's f32 metric lies at or below T, that candidate is the f64 arg-min of the
reference's metric (first minimum in scan order).  Checked on clustered, lattice (exact ties), near-duplicate and far-from-origin
scenes — the cases where f32 could mis-rank — and the measured worst |m32 - m| is compared with e(D).  No GPU involved: this
pins the error analysis itself; tests/test_gpu_search_exactness.py pins the kernel."""
import numpy as np
import pytest

U = 2.0 ** -24


def _f32(x):
    return np.asarray(x, dtype=np.float32)


def _fma(a, b, c):  # fmaf: exact product (24 x 24 bits fit a double), one f64 add, rounded to f32
    return _f32(a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64))


def _trunc_key(x, vs):
    return np.trunc(x / vs).astype(np.int64)


def _model(pts, q, vs, th):
    """For every query: f64 metric of all candidates in its 27 voxels (reference order irrelevant here: candidates are given as one
    set per query), the f32 metric as the kernel computes it, and the band decision."""
    s, smax = min(1.0, th), max(1.0, th)
    smin32 = np.float32(s * (1 - 1e-5)); inv_smin32 = np.float32((1 + 1e-5) / s)
    err_scale = np.float32(smax * (1 + 1e-5))
    err_a, err_b, err_c = np.float32(64 * U * vs), np.float32(12 * U), np.float32(1e-11 * vs * vs)
    e = lambda D: err_scale * (err_a * np.sqrt(_f32(D)) + err_b * _f32(D)) + err_c
    vs32, th32 = np.float32(vs), np.float32(th)

    kq = _trunc_key(q[:, :3], vs)
    kp = _trunc_key(pts[:, :3], vs)
    # record: f32(coord - key * vs) ; label exact
    rec = _f32(pts[:, :3] - kp * vs)
    bq = _f32(q[:, :3] - kq * vs)
    out = []
    for i in range(len(q)):
        off = kp - kq[i]
        sel = np.flatnonzero((np.abs(off) <= 1).all(1))
        if len(sel) == 0:
            continue
        # exact metric (reference arithmetic: f64)
        d = ((pts[sel, :3] - q[i, :3]) ** 2).sum(1)
        ln, lq = np.trunc(pts[sel, 3]), np.trunc(q[i, 3])
        compat = (ln == lq) | (np.trunc(pts[sel, 3] * q[i, 3]) == 0)
        m = np.where(compat, d * th, d)
        # f32 metric: query relative to the candidate's voxel origin, then dx, dy, dz and the fma chain
        rel = _f32(bq[i][None, :] - _f32(off[sel].astype(np.float32) * vs32))
        dx, dy, dz = (rec[sel, 0] - rel[:, 0]), (rec[sel, 1] - rel[:, 1]), (rec[sel, 2] - rel[:, 2])
        D32 = _fma(dz, dz, _fma(dy, dy, _f32(dx * dx)))
        m32 = np.where(compat, _f32(D32 * th32), D32)
        gm = m32.min()
        e1 = e(gm * inv_smin32)
        T = gm + np.float32(1.01) * (e1 + e((gm + e1) * inv_smin32))
        inside = np.flatnonzero(m32 <= T)
        out.append((m, m32, d, inside, e(_f32(d))))
    return out


def _scene(kind, rng):
    if kind == "clustered":
        c = rng.uniform(-6, 6, (10, 3))
        pts = c[rng.integers(0, 10, 20000)] + rng.normal(0, 0.5, (20000, 3)) * [1, 1, 0.05]
        q = pts[rng.integers(0, len(pts), 1500)] + rng.normal(0, 0.3, (1500, 3))
    elif kind == "lattice":
        ax = np.arange(-12, 12) * 0.25
        X, Y, Z = np.meshgrid(ax, ax, ax[8:16], indexing="ij")
        pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], 1)
        q = pts[rng.integers(0, len(pts), 1500)].copy()
        q[:700] += 0.125
    elif kind == "near_duplicates":
        base = rng.uniform(-4, 4, (4000, 3))
        pts = np.concatenate([base, base + rng.normal(0, 3e-8, base.shape), base + rng.normal(0, 2e-7, base.shape)])
        q = base[rng.integers(0, len(base), 1500)] + rng.normal(0, 0.2, (1500, 3))
    else:  # far from the origin: 2e5 m, where a world-frame f32 would have 1.5 cm resolution
        pts = rng.uniform(-5, 5, (20000, 3)) + [1.3e5, 2.0e5, -900.0]
        q = pts[rng.integers(0, len(pts), 1500)] + rng.normal(0, 0.4, (1500, 3))
    labels = rng.choice([0, 40, 50, 80, 81], len(pts)).astype(float)
    return np.c_[pts, labels], np.c_[q, rng.choice([0, 40, 50, 81, 10], len(q)).astype(float)]


@pytest.mark.parametrize("kind", ["clustered", "lattice", "near_duplicates", "far"])
@pytest.mark.parametrize("vs,th", [(0.8, 0.4), (0.3, 1.0), (1.7, 2.5), (0.8, 0.05)])
def test_unique_within_band_is_the_exact_argmin(kind, vs, th):
    import zlib
    rng = np.random.default_rng(zlib.crc32(f"{kind} {vs} {th}".encode()))
    pts, q = _scene(kind, rng)
    decided = ambiguous = 0
    worst_ratio = 0.0
    for m, m32, d, inside, e_d in _model(pts, q, vs, th):
        err = np.abs(m32.astype(np.float64) - m)
        worst_ratio = max(worst_ratio, float((err / e_d.astype(np.float64)).max()))
        if len(inside) == 1:
            decided += 1
            j = inside[0]
            assert m[j] == m.min() and (m == m.min()).sum() == 1, (kind, vs, th)  # the f64 arg-min, and no exact tie hidden
        else:
            ambiguous += 1  # goes to the f64 path on the device
    assert decided + ambiguous > 1000
    assert worst_ratio < 0.5  # the error model e(D) has at least 2x slack over everything observed
    if kind in ("clustered", "far"):
        assert ambiguous < 0.02 * (decided + ambiguous)  # the fast path serves nearly everything on ordinary scenes
    if kind == "lattice":
        assert ambiguous > 300  # exact ties are never decided in f32
