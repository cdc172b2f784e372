"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/liboracle.so (the CPU restatement of the
reference hot path).  Imported only by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs; never by the sage_icp_b200 package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.environ.get("SAGE_ORACLE_LIB") or os.path.join(_HERE, "liboracle.so")  # override: sanitizer builds

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("oracle_capi.cpp", "sage_oracle.hpp", "robin_table.hpp", "se3.hpp", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if os.environ.get("SAGE_ORACLE_LIB"):
        return _LIB_PATH
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_create.restype = C.c_void_p
        L.orc_map_create.restype = C.c_void_p
        L.orc_pipeline_map.restype = C.c_void_p
        for f in ("orc_preprocess", "orc_preprocess_dynamic", "orc_preprocess_dynamic_ordered", "orc_voxel_downsample", "orc_map_num_voxels", "orc_map_bucket_count", "orc_map_num_points",
                  "orc_map_pointcloud", "orc_map_dump", "orc_map_get_correspondences", "orc_last_source",
                  "orc_last_frame_downsample", "orc_num_poses", "orc_local_map", "orc_robin_order", "orc_robin_replay", "orc_voxelize", "orc_deskew"):
            getattr(L, f).restype = C.c_size_t
        for f in ("orc_get_adaptive_threshold", "orc_last_sigma", "orc_rotation_angle", "orc_occ_overlap"):
            getattr(L, f).restype = C.c_double
        L.orc_voxel_hash.restype = C.c_uint32
        _lib = L
    return _lib


def _d(a: np.ndarray):
    return a.ctypes.data_as(_dp)


def _c64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


# ---- SE(3) -----------------------------------------------------------------------------------
def se3_exp(xi) -> np.ndarray:
    xi = _c64(xi); out = np.empty(7)
    lib().orc_se3_exp(_d(xi), _d(out)); return out


def se3_log(pose) -> np.ndarray:
    pose = _c64(pose); out = np.empty(6)
    lib().orc_se3_log(_d(pose), _d(out)); return out


def se3_mul(a, b) -> np.ndarray:
    a, b = _c64(a), _c64(b); out = np.empty(7)
    lib().orc_se3_mul(_d(a), _d(b), _d(out)); return out


def se3_inverse(a) -> np.ndarray:
    a = _c64(a); out = np.empty(7)
    lib().orc_se3_inverse(_d(a), _d(out)); return out


def se3_act(a, p) -> np.ndarray:
    a, p = _c64(a), _c64(p); out = np.empty(3)
    lib().orc_se3_act(_d(a), _d(p), _d(out)); return out


def rotation_angle(pose) -> float:
    pose = _c64(pose)
    return float(lib().orc_rotation_angle(_d(pose)))


def ldlt6_solve(A, b) -> np.ndarray:
    A, b = _c64(A), _c64(b); x = np.empty(6)
    lib().orc_ldlt6_solve(_d(A), _d(b), _d(x)); return x


def voxel_hash(x: int, y: int, z: int) -> int:
    return int(lib().orc_voxel_hash(C.c_int32(x), C.c_int32(y), C.c_int32(z)))


def robin_order(keys: np.ndarray) -> Tuple[np.ndarray, int]:
    keys = np.ascontiguousarray(keys, dtype=np.int32)
    order = np.empty(len(keys), dtype=np.int64)
    bc = lib().orc_robin_order(keys.ctypes.data_as(_ip), C.c_size_t(len(keys)), order.ctypes.data_as(_lp))
    return order, int(bc)


# ---- core free functions ---------------------------------------------------------------------
def robin_replay(ops: np.ndarray) -> Tuple[np.ndarray, int]:
    """ops (n, 4) int32 = kind (0 insert, 1 erase, 2 sweep-erase keys with x < ops.x while iterating), x, y, z.
    Returns (final iteration order as the op indices that inserted the surviving keys, bucket_count)."""
    ops = np.ascontiguousarray(ops, np.int32); order = np.empty(len(ops), np.int64); bc = C.c_int64()
    n = lib().orc_robin_replay(ops.ctypes.data_as(_ip), C.c_size_t(len(ops)), order.ctypes.data_as(_lp), C.byref(bc))
    return order[:n].copy(), int(bc.value)


def preprocess(pts, max_range, min_range, label_max_range) -> np.ndarray:
    pts = _c64(pts); out = np.empty_like(pts)
    n = lib().orc_preprocess(_d(pts), C.c_size_t(len(pts)), C.c_double(max_range), C.c_double(min_range),
                             C.c_double(label_max_range), _d(out), C.c_size_t(len(pts)))
    return out[:n].copy()


def preprocess_dynamic(cfg, pts, cluster_order: bool = False) -> np.ndarray:
    """Preprocess with the dynamic-vehicle filter on (core/Preprocessing.cpp:95-172).  cluster_order: re-admitted vehicle points
    cluster by cluster in PCL's published order (as the reference emits them) instead of input order (what the CUDA path does)."""
    pod = cfg.to_pod(); pts = _c64(pts); out = np.empty_like(pts)
    n = lib().orc_preprocess_dynamic_ordered(C.byref(pod), _d(pts), C.c_size_t(len(pts)), int(cluster_order), _d(out), C.c_size_t(len(pts)))
    return out[:n].copy()


def voxel_downsample(cfg, pts, vox_scale) -> np.ndarray:
    pod = cfg.to_pod(); pts = _c64(pts); out = np.empty_like(pts)
    n = lib().orc_voxel_downsample(C.byref(pod), _d(pts), C.c_size_t(len(pts)), C.c_double(vox_scale), _d(out), C.c_size_t(len(pts)))
    return out[:n].copy()


def align_clouds(src, tgt, th, threads=1):
    src, tgt = _c64(src), _c64(tgt)
    JTJ, JTr, x, est = np.empty((6, 6)), np.empty(6), np.empty(6), np.empty(7)
    lib().orc_align_clouds(_d(src), _d(tgt), C.c_size_t(len(src)), C.c_double(th), C.c_int(threads), _d(JTJ), _d(JTr), _d(x), _d(est))
    return JTJ, JTr, x, est


def grid_map(pts, bounds, rows: int, cols: int) -> np.ndarray:
    """utils::EigenToGridMap (ros/ros2/Utils.hpp:220-242)."""
    pts = _c64(pts); b = _c64(np.asarray(bounds, float).reshape(6)); out = np.zeros((rows, cols), np.int32)
    lib().orc_grid_map(_d(pts), C.c_size_t(len(pts)), _d(b), rows, cols, out.ctypes.data_as(_ip))
    return out


def occ_overlap(occ_s, occ_t) -> float:
    """utils::compute_occ_overlap (ros/ros2/Utils.hpp:244-258)."""
    a, b = np.ascontiguousarray(occ_s, np.int32), np.ascontiguousarray(occ_t, np.int32)
    return float(lib().orc_occ_overlap(a.ctypes.data_as(_ip), b.ctypes.data_as(_ip), C.c_size_t(a.size)))


def deskew(frame, ts, start, finish) -> np.ndarray:
    frame, ts, start, finish = _c64(frame), _c64(ts), _c64(start), _c64(finish)
    out = np.empty_like(frame)
    lib().orc_deskew(_d(frame), _d(ts), C.c_size_t(len(frame)), _d(start), _d(finish), _d(out))
    return out


class OracleMap:
    """VoxelHashMap restatement (core/VoxelHashMap.hpp)."""

    def __init__(self, voxel_size, max_distance, basic, critical, basic_labels, evict_faithful=True, _borrow=None):
        self._own = _borrow is None
        self.voxel_size = voxel_size
        self.stride = basic + critical
        if _borrow is None:
            lab = (C.c_int32 * max(1, len(basic_labels)))(*basic_labels)
            self.h = C.c_void_p(lib().orc_map_create(C.c_double(voxel_size), C.c_double(max_distance), basic, critical, lab,
                                                     len(basic_labels), int(evict_faithful)))
        else:
            self.h = C.c_void_p(_borrow)

    def __del__(self):
        if getattr(self, "_own", False) and self.h:
            lib().orc_map_destroy(self.h); self.h = None

    def clear(self): lib().orc_map_clear(self.h)
    def num_voxels(self) -> int: return int(lib().orc_map_num_voxels(self.h))
    def bucket_count(self) -> int: return int(lib().orc_map_bucket_count(self.h))
    def num_points(self) -> int: return int(lib().orc_map_num_points(self.h))

    def add_points(self, pts):
        pts = _c64(pts); lib().orc_map_add_points(self.h, _d(pts), C.c_size_t(len(pts)))

    def remove_far(self, origin):
        o = _c64(origin); lib().orc_map_remove_far(self.h, _d(o))

    def update(self, pts, pose):
        pts, pose = _c64(pts), _c64(pose); lib().orc_map_update(self.h, _d(pts), C.c_size_t(len(pts)), _d(pose))

    def pointcloud(self) -> np.ndarray:
        n = self.num_points(); out = np.empty((n, 4))
        lib().orc_map_pointcloud(self.h, _d(out), C.c_size_t(n)); return out

    def dump(self):
        """(keys V x3 int32, counts V int32, points V x stride x4 f64) in robin iteration order."""
        v = self.num_voxels()
        keys = np.zeros((v, 3), np.int32); counts = np.zeros(v, np.int32); pts = np.zeros((v, self.stride, 4))
        lib().orc_map_dump(self.h, keys.ctypes.data_as(_ip), counts.ctypes.data_as(_ip), _d(pts), self.stride, C.c_size_t(v))
        return keys, counts, pts

    def get_correspondences(self, pts, max_dist, th, threads=1):
        pts = _c64(pts); n = len(pts)
        src, tgt, q = np.empty((n, 4)), np.empty((n, 4)), np.empty(n, np.int64)
        k = lib().orc_map_get_correspondences(self.h, _d(pts), C.c_size_t(n), C.c_double(max_dist), C.c_double(th), threads,
                                              _d(src), _d(tgt), q.ctypes.data_as(_lp))
        return src[:k].copy(), tgt[:k].copy(), q[:k].copy()

    def nn_stats(self, pts) -> Tuple[int, int]:
        pts = _c64(pts); o, c = C.c_uint64(), C.c_uint64()
        lib().orc_map_nn_stats(self.h, _d(pts), C.c_size_t(len(pts)), C.byref(o), C.byref(c))
        return int(o.value), int(c.value)

    def register_frame_core(self, frame, guess, max_dist, kernel, sem_th, threads=1, max_iters=500, est_th=1e-4):
        frame, guess = _c64(frame), _c64(guess); out = np.empty(7)
        it = lib().orc_register_frame_core(self.h, _d(frame), C.c_size_t(len(frame)), _d(guess), C.c_double(max_dist),
                                           C.c_double(kernel), C.c_double(sem_th), threads, max_iters, C.c_double(est_th), _d(out))
        return out, int(it)


class OraclePipeline:
    """sageICP restatement (pipeline/sageICP.hpp)."""

    def __init__(self, cfg, threads: int = 1, evict_faithful: bool = True):
        self.cfg = cfg
        self._pod = cfg.to_pod()
        self.h = C.c_void_p(lib().orc_create(C.byref(self._pod), threads, int(evict_faithful)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h); self.h = None

    def reset(self): lib().orc_reset(self.h)

    def set_dynamic_cluster_order(self, on: bool):
        """Dynamic-vehicle filter: emit the re-admitted vehicle points cluster by cluster (the reference's order) instead of input order."""
        lib().orc_set_dynamic_cluster_order(self.h, int(bool(on)))

    def register_frame(self, pts, timestamps: Optional[np.ndarray] = None):
        pts = _c64(pts); pose = np.empty(7); ti, ta = C.c_double(), C.c_double()
        ts = None if timestamps is None else _d(_c64(timestamps))
        lib().orc_register_frame(self.h, _d(pts), C.c_size_t(len(pts)), ts, _d(pose), C.byref(ti), C.byref(ta))
        return pose, ti.value, ta.value

    def voxelize(self, pts):
        pts = _c64(pts); s, d = np.empty_like(pts), np.empty_like(pts); ns, nd = C.c_size_t(), C.c_size_t()
        lib().orc_voxelize(self.h, _d(pts), C.c_size_t(len(pts)), _d(s), C.byref(ns), _d(d), C.byref(nd))
        return s[:ns.value].copy(), d[:nd.value].copy()

    def _cloud(self, fn, cap_fn=None):
        n = fn(self.h, None, C.c_size_t(0)); out = np.empty((n, 4))
        fn(self.h, _d(out), C.c_size_t(n)); return out

    def last_source(self): return self._cloud(lib().orc_last_source)
    def last_frame_downsample(self): return self._cloud(lib().orc_last_frame_downsample)
    def local_map(self): return self._cloud(lib().orc_local_map)
    def last_iterations(self) -> int: return int(lib().orc_last_iterations(self.h))
    def last_sigma(self) -> float: return float(lib().orc_last_sigma(self.h))
    def adaptive_threshold(self) -> float: return float(lib().orc_get_adaptive_threshold(self.h))
    def has_moved(self) -> bool: return bool(lib().orc_has_moved(self.h))

    def prediction_model(self):
        out = np.empty(7); lib().orc_get_prediction_model(self.h, _d(out)); return out

    def poses(self) -> np.ndarray:
        n = int(lib().orc_num_poses(self.h)); out = np.empty((n, 7))
        for i in range(n):
            lib().orc_get_pose(self.h, C.c_size_t(i), _d(out[i]))
        return out

    def map(self) -> OracleMap:
        c = self.cfg
        return OracleMap(c.voxel_size_map, c.local_map_range, c.basic_points_per_voxel, c.critical_points_per_voxel,
                         c.basic_parts_labels, _borrow=lib().orc_pipeline_map(self.h))


def max_threads() -> int:
    return int(lib().orc_max_threads())
