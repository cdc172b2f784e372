"""Seeded synthetic inputs for the SAGE-ICP hot path (harness helper used by tests/ and bench.py).

Shapes follow SURVEY.md §8d: an HDL-64-like scan (64 beams, +2 deg .. -24.8 deg, x 1875 azimuths = 120 000
rays; 128 x 3907 ~ 500 k for the dense config) cast from 1.73 m into a procedural labelled street, Gaussian
range noise sigma = 0.02 m, xyz rounded to float32 and widened (the real input is f32 xyz + u8 label,
ros/ros2/Utils.hpp:161-180).  Labels are SemanticKITTI ids (ros/launch/semantic-kitti.yaml).  Labels 30 and 252
belong to no voxel group, so VoxelDownsample's "dropped" path is exercised (core/Preprocessing.cpp:69).

Everything is numpy, deterministic for a given (seed, pose, shape).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

SENSOR_HEIGHT = 1.73
WALL_Y = 10.0
GROUND_Y = 13.5
NO_HIT_RANGE = 150.0  # rays that hit nothing come back beyond max_range and are cropped by Preprocess


def _hash01(ix: np.ndarray, salt: int) -> np.ndarray:
    """Cheap deterministic per-cell hash in [0,1)."""
    h = (ix.astype(np.int64) * 2654435761 + salt * 40503) & 0xFFFFFFFF
    h ^= h >> 15
    h = (h * 2246822519) & 0xFFFFFFFF
    h ^= h >> 13
    return h.astype(np.float64) / 4294967296.0


@dataclass
class Prim:
    kind: str  # "cyl" | "box" | "sph"
    c: Tuple[float, float, float]  # centre (cyl: axis base centre x,y and z0)
    s: Tuple[float, float, float]  # cyl: (r, h, _) ; box: half sizes ; sph: (r,_,_)
    label: int


def street_prims(x_lo: float, x_hi: float) -> List[Prim]:
    """Objects of the procedural street whose anchor x lies in [x_lo, x_hi]."""
    prims: List[Prim] = []

    def h(k: int, sgn: int, salt: int) -> float:
        return float(_hash01(np.array([k * 2 + sgn]), salt)[0])

    for side in (-1.0, 1.0):
        sgn = 0 if side < 0 else 1
        # facades: one building (50) or fence (51) box per 4 m cell, set back by a per-cell depth, plus a
        # protruding pillar per building cell, so the street is not translation-invariant along x
        for k in range(math.floor(x_lo / 4.0), math.ceil(x_hi / 4.0) + 1):
            h1, h2, h3 = h(k, sgn, 3), h(k, sgn, 5), h(k, sgn, 9)
            yw = WALL_Y + 2.5 * h2
            if h1 < 0.25:
                prims.append(Prim("box", (4.0 * k + 2.0, side * (yw + 0.1), 1.0), (2.0, 0.1, 1.0), 51))
            else:
                hgt = 5.0 + 7.0 * h1
                prims.append(Prim("box", (4.0 * k + 2.0, side * (yw + 4.0), hgt / 2), (2.0, 4.0, hgt / 2), 50))
                prims.append(Prim("box", (4.0 * k + 0.5 + 3.0 * h3, side * (yw - 0.3), 1.6), (0.3, 0.3, 1.6), 50))
        # poles (80) + traffic signs (81) every 10 m
        for k in range(math.floor(x_lo / 10.0), math.ceil(x_hi / 10.0) + 1):
            x = 10.0 * k + 4.0 * h(k, sgn, 15)
            prims.append(Prim("cyl", (x, side * 6.5, 0.0), (0.12, 7.0, 0.0), 80))
            if h(k, sgn, 19) < 0.5:
                prims.append(Prim("box", (x + 1.5, side * 5.6, 2.4), (0.05, 0.4, 0.4), 81))
        # trunks (71) + vegetation blobs (70) every 7 m
        for k in range(math.floor(x_lo / 7.0), math.ceil(x_hi / 7.0) + 1):
            j = h(k, sgn, 7)
            x = 7.0 * k + 4.0 * j
            prims.append(Prim("cyl", (x, side * 8.0, 0.0), (0.25, 3.2, 0.0), 71))
            prims.append(Prim("sph", (x, side * 8.0, 4.4), (1.2 + 0.8 * j, 0.0, 0.0), 70))
        # parked cars (10) / moving cars (252, in no voxel group) every 8 m when the cell hash says so
        for k in range(math.floor(x_lo / 8.0), math.ceil(x_hi / 8.0) + 1):
            j = h(k, sgn, 11)
            if j < 0.5:
                x = 8.0 * k + 3.0 * j
                prims.append(Prim("box", (x, side * 3.4, 0.78), (2.1, 0.9, 0.75), 10 if j > 0.05 else 252))
        # pedestrians (30, dropped) and other-objects (99) on the sidewalk every 12 m
        for k in range(math.floor(x_lo / 12.0), math.ceil(x_hi / 12.0) + 1):
            j = h(k, sgn, 13)
            x = 12.0 * k + 6.0 * j
            if j < 0.4:
                prims.append(Prim("cyl", (x, side * 5.0, 0.0), (0.3, 1.75, 0.0), 30))
            else:
                prims.append(Prim("box", (x, side * 7.2, 0.5), (0.6, 0.4, 0.5), 99))
    return prims


def _ground_label(y: np.ndarray, x: np.ndarray) -> np.ndarray:
    ay = np.abs(y)
    lab = np.full(y.shape, 72, dtype=np.int32)  # terrain
    lab[ay < 6.0] = 48  # sidewalk
    lab[ay < 4.0] = 40  # road
    park = (ay >= 2.4) & (ay < 4.0) & (_hash01(np.floor(x / 12.0).astype(np.int64), 17) < 0.3)
    lab[park] = 44  # parking
    return lab


def ray_grid(n_beams: int = 64, n_az: int = 1875) -> Tuple[np.ndarray, np.ndarray]:
    elev = np.deg2rad(np.linspace(2.0, -24.8, n_beams))
    az = np.linspace(-math.pi, math.pi, n_az, endpoint=False)
    return elev, az


def make_scan(seed: int, pose_xyyaw: Tuple[float, float, float] = (0.0, 0.0, 0.0), n_beams: int = 64,
              n_az: int = 1875, noise: float = 0.02, label_flip: float = 0.02) -> np.ndarray:
    """One labelled scan in the SENSOR frame: float64 array (n_beams*n_az, 4) = x, y, z, label."""
    px, py, yaw = pose_xyyaw
    elev, az = ray_grid(n_beams, n_az)
    ce, se = np.cos(elev)[:, None], np.sin(elev)[:, None]
    # ray directions in the world frame (sensor yawed about z)
    dx = ce * np.cos(az + yaw)[None, :]
    dy = ce * np.sin(az + yaw)[None, :]
    dz = np.broadcast_to(se, dx.shape).copy()
    ox, oy, oz = px, py, SENSOR_HEIGHT

    t_best = np.full(dx.shape, np.inf)
    lab = np.zeros(dx.shape, dtype=np.int32)

    # ground z = 0
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.where(dz < -1e-9, -oz / dz, np.inf)
    tf = np.where(np.isfinite(t), t, 0.0)
    hx, hy = ox + tf * dx, oy + tf * dy
    ok = np.isfinite(t) & (np.abs(hy) < GROUND_Y)
    t_best = np.where(ok, t, t_best)
    gl = _ground_label(np.where(ok, hy, 0.0), np.where(ok, hx, 0.0))
    lab = np.where(ok, gl, lab)

    # objects: only test the azimuth window that can see each primitive
    daz = 2.0 * math.pi / n_az
    for p in street_prims(px - 105.0, px + 105.0):
        cx, cy, cz = p.c
        rel = math.hypot(cx - ox, cy - oy)
        rad = {"cyl": p.s[0], "box": math.hypot(p.s[0], p.s[1]), "sph": p.s[0]}[p.kind]
        if rel <= rad + 1e-6 or rel > 110.0:
            continue
        bearing = math.atan2(cy - oy, cx - ox) - yaw
        half = math.asin(min(1.0, rad / rel)) + 2 * daz
        i0 = int(math.floor((bearing - half + math.pi) / daz))
        i1 = int(math.ceil((bearing + half + math.pi) / daz))
        cols = np.arange(i0, i1 + 1) % n_az
        ddx, ddy, ddz = dx[:, cols], dy[:, cols], dz[:, cols]
        if p.kind == "cyl":
            r, h, _ = p.s
            a = ddx * ddx + ddy * ddy
            fx, fy = ox - cx, oy - cy
            b = 2.0 * (fx * ddx + fy * ddy)
            c = fx * fx + fy * fy - r * r
            disc = b * b - 4 * a * c
            with np.errstate(invalid="ignore", divide="ignore"):
                t = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), np.inf)
            hz = oz + np.where(np.isfinite(t), t, 0.0) * ddz
            ok = np.isfinite(t) & (t > 0) & (hz >= cz) & (hz <= cz + h)
        elif p.kind == "sph":
            r = p.s[0]
            fx, fy, fz = ox - cx, oy - cy, oz - cz
            b = 2.0 * (fx * ddx + fy * ddy + fz * ddz)
            c = fx * fx + fy * fy + fz * fz - r * r
            disc = b * b - 4 * c
            with np.errstate(invalid="ignore"):
                t = np.where(disc > 0, (-b - np.sqrt(np.maximum(disc, 0))) / 2.0, np.inf)
            ok = np.isfinite(t) & (t > 0)
        else:  # axis-aligned box, slab method
            lo = np.array([cx - p.s[0], cy - p.s[1], cz - p.s[2]])
            hi = np.array([cx + p.s[0], cy + p.s[1], cz + p.s[2]])
            o = (ox, oy, oz)
            tmin = np.full(ddx.shape, -np.inf)
            tmax = np.full(ddx.shape, np.inf)
            for ax, dd in enumerate((ddx, ddy, ddz)):
                with np.errstate(divide="ignore", invalid="ignore"):
                    inv = 1.0 / np.where(np.abs(dd) < 1e-12, 1e-12, dd)
                ta, tb = (lo[ax] - o[ax]) * inv, (hi[ax] - o[ax]) * inv
                tmin = np.maximum(tmin, np.minimum(ta, tb))
                tmax = np.minimum(tmax, np.maximum(ta, tb))
            ok = (tmax >= tmin) & (tmin > 0)
            t = np.where(ok, tmin, np.inf)
        sub = t_best[:, cols]
        ok = ok & (t < sub)
        if ok.any():
            sub = np.where(ok, t, sub)
            t_best[:, cols] = sub
            sl = lab[:, cols]
            lab[:, cols] = np.where(ok, p.label, sl)

    rng = np.random.default_rng(seed)
    hit = np.isfinite(t_best) & (t_best < NO_HIT_RANGE)
    rngs = np.where(hit, t_best, NO_HIT_RANGE) + rng.normal(0.0, noise, size=t_best.shape)
    lab = np.where(hit, lab, 0)
    if label_flip > 0:
        flip = rng.random(size=lab.shape) < label_flip
        pool = np.array([0, 10, 30, 40, 44, 48, 50, 51, 70, 71, 72, 80, 81, 99, 252], dtype=np.int32)
        lab = np.where(flip, pool[rng.integers(0, len(pool), size=lab.shape)], lab)
    # back to the sensor frame: direction in sensor frame is (ce*cos az, ce*sin az, se)
    sx = rngs * (ce * np.cos(az)[None, :])
    sy = rngs * (ce * np.sin(az)[None, :])
    sz = rngs * np.broadcast_to(se, rngs.shape)
    out = np.empty((n_beams * n_az, 4), dtype=np.float64)
    # azimuth-major (a spinning lidar reports column by column)
    out[:, 0] = sx.T.reshape(-1).astype(np.float32)
    out[:, 1] = sy.T.reshape(-1).astype(np.float32)
    out[:, 2] = sz.T.reshape(-1).astype(np.float32)
    out[:, 3] = lab.T.reshape(-1).astype(np.float64)
    return out


def trajectory(n_frames: int, step: float = 1.0, yaw_amp_deg: float = 0.5, period: int = 60) -> np.ndarray:
    """KITTI-like motion: `step` m/frame forward with sinusoidal yaw; returns (n,3) x,y,yaw."""
    poses = np.zeros((n_frames, 3))
    x = y = yaw = 0.0
    for i in range(n_frames):
        poses[i] = (x, y, yaw)
        v = step * min(1.0, (i + 1) / 8.0)  # pull away from rest (the constant-velocity model starts at 0)
        yaw += math.radians(yaw_amp_deg) * math.sin(2 * math.pi * i / period)
        x += v * math.cos(yaw)
        y += v * math.sin(yaw)
    return poses


def pose7_from_xyyaw(p: Tuple[float, float, float], z: float = SENSOR_HEIGHT) -> np.ndarray:
    """(x, y, yaw) -> [tx,ty,tz,qx,qy,qz,qw] of the sensor in the world."""
    x, y, yaw = p
    return np.array([x, y, z, 0.0, 0.0, math.sin(yaw / 2), math.cos(yaw / 2)])


def sample_street_map(n_points: int, seed: int, x_lo: float, x_hi: float, noise: float = 0.02) -> np.ndarray:
    """World-frame labelled surface samples of the street between x_lo and x_hi (for pre-built maps,
    BASELINE configs 1/2/5).  float64 (n,4); z is height above the ground, so a sensor pose from
    pose7_from_xyyaw() sees these points where make_scan() would put them."""
    rng = np.random.default_rng(seed)
    n_ground = int(n_points * 0.4)
    n_obj = n_points - n_ground
    out = np.empty((n_points, 4))
    gx = rng.uniform(x_lo, x_hi, n_ground)
    gy = rng.uniform(-GROUND_Y, GROUND_Y, n_ground)
    out[:n_ground, 0], out[:n_ground, 1] = gx, gy
    out[:n_ground, 2] = rng.normal(0.0, noise, n_ground)
    out[:n_ground, 3] = _ground_label(gy, gx)

    prims = [p for p in street_prims(x_lo, x_hi) if x_lo <= p.c[0] <= x_hi]
    kind = np.array([{"cyl": 0, "sph": 1, "box": 2}[p.kind] for p in prims])
    Cc = np.array([p.c for p in prims], dtype=np.float64).reshape(-1, 3)
    Ss = np.array([p.s for p in prims], dtype=np.float64).reshape(-1, 3)
    Ll = np.array([p.label for p in prims], dtype=np.float64)
    area = np.where(kind == 0, 2 * math.pi * Ss[:, 0] * Ss[:, 1],
                    np.where(kind == 1, 4 * math.pi * Ss[:, 0] ** 2,
                             8 * (Ss[:, 0] * Ss[:, 1] + Ss[:, 1] * Ss[:, 2] + Ss[:, 0] * Ss[:, 2])))
    # small objects get a floor so poles/signs are well represented
    w = np.maximum(area, 6.0)
    cdf = np.cumsum(w) / w.sum()
    which = np.minimum(np.searchsorted(cdf, rng.random(n_obj)), len(prims) - 1)
    u, v, w3 = rng.random(n_obj), rng.random(n_obj), rng.random(n_obj)
    k = kind[which]
    c, sz = Cc[which], Ss[which]
    pts = np.empty((n_obj, 3))
    m = k == 0
    ang = 2 * math.pi * u[m]
    pts[m, 0] = c[m, 0] + sz[m, 0] * np.cos(ang)
    pts[m, 1] = c[m, 1] + sz[m, 0] * np.sin(ang)
    pts[m, 2] = c[m, 2] + sz[m, 1] * v[m]
    m = k == 1
    ang, ct = 2 * math.pi * u[m], 2 * v[m] - 1
    st = np.sqrt(1 - ct * ct)
    pts[m, 0] = c[m, 0] + sz[m, 0] * st * np.cos(ang)
    pts[m, 1] = c[m, 1] + sz[m, 0] * st * np.sin(ang)
    pts[m, 2] = c[m, 2] + sz[m, 0] * ct
    m = np.nonzero(k == 2)[0]
    loc = np.stack([(2 * u[m] - 1) * sz[m, 0], (2 * v[m] - 1) * sz[m, 1], (2 * w3[m] - 1) * sz[m, 2]], axis=1)
    # pick a face with probability ~ its area: x-faces ~ sy*sz, y-faces ~ sx*sz, z-faces ~ sx*sy
    fa = np.stack([sz[m, 1] * sz[m, 2], sz[m, 0] * sz[m, 2], sz[m, 0] * sz[m, 1]], axis=1)
    fc = np.cumsum(fa, axis=1) / fa.sum(axis=1, keepdims=True)
    r = rng.random(len(m))[:, None]
    face = np.minimum((r > fc).sum(axis=1), 2)
    sg = np.where(rng.random(len(m)) < 0.5, -1.0, 1.0)
    loc[np.arange(len(m)), face] = sg * sz[m][np.arange(len(m)), face]
    pts[m] = c[m] + loc
    out[n_ground:, :3] = pts + rng.normal(0.0, noise, size=(n_obj, 3))
    out[n_ground:, 3] = Ll[which]
    keep = out[:, 2] > -0.2  # nothing below the ground
    out = out[keep]
    return out[rng.permutation(len(out))]


def world_to_sensor(points_w: np.ndarray, pose7: np.ndarray) -> np.ndarray:
    """Inverse rigid transform of world points into the sensor frame of `pose7` (yaw-only poses)."""
    yaw = 2.0 * math.atan2(pose7[5], pose7[6])
    c, s = math.cos(yaw), math.sin(yaw)
    d = points_w[:, :3] - pose7[:3]
    out = points_w.copy()
    out[:, 0] = c * d[:, 0] + s * d[:, 1]
    out[:, 1] = -s * d[:, 0] + c * d[:, 1]
    out[:, 2] = d[:, 2]
    return out
