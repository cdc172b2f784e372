"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes binding of oracle/_ref/libsage_ref.so: the REFERENCE's own hot-path sources
(cpp/sage_icp/core/*.cpp, pipeline/sageICP.cpp), compiled unmodified from /root/reference against the stand-in headers of
oracle/shim/ (oracle/ref_capi.cpp, oracle/Makefile target `ref`).  Used by tests/test_reference_build.py to check the oracle's
restatement against the reference's own code.  Never imported by the sage_icp_b200 package, bench.py or smoke()."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SAGE_REF_LIB") or os.path.join(_HERE, "_ref", "libsage_ref.so")  # override: experiments with other compiler flags
REFERENCE_ROOT = "/root/reference/cpp"

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lib = None


def available() -> bool:
    """True if the reference build exists (built earlier and shipped) or can be built here (/root/reference present)."""
    return os.path.exists(LIB_PATH) or os.path.isdir(os.path.join(REFERENCE_ROOT, "sage_icp", "core"))


def build() -> str:
    if os.path.isdir(os.path.join(REFERENCE_ROOT, "sage_icp", "core")):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("oracle/_ref/libsage_ref.so is missing and /root/reference is not here to build it from")
    return LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.ref_create.restype = C.c_void_p
        L.ref_map_create.restype = C.c_void_p
        for f in ("ref_preprocess", "ref_voxel_downsample", "ref_deskew", "ref_map_num_voxels", "ref_map_pointcloud", "ref_map_dump",
                  "ref_map_get_correspondences", "ref_voxelize", "ref_last_source", "ref_num_poses", "ref_local_map"):
            getattr(L, f).restype = C.c_size_t
        L.ref_get_adaptive_threshold.restype = C.c_double
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _c64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def preprocess(cfg, pts) -> np.ndarray:
    """sage_icp::Preprocess with the config's range / dynamic-filter parameters (core/Preprocessing.cpp:86-189)."""
    pod = cfg.to_pod(); pts = _c64(pts); out = np.empty_like(pts)
    n = lib().ref_preprocess(C.byref(pod), _d(pts), C.c_size_t(len(pts)), _d(out), C.c_size_t(len(pts)))
    return out[:n].copy()


def voxel_downsample(cfg, pts, vox_scale) -> np.ndarray:
    pod = cfg.to_pod(); pts = _c64(pts); out = np.empty_like(pts)
    n = lib().ref_voxel_downsample(C.byref(pod), _d(pts), C.c_size_t(len(pts)), C.c_double(vox_scale), _d(out), C.c_size_t(len(pts)))
    return out[:n].copy()


def deskew(frame, ts, start, finish) -> np.ndarray:
    frame, ts, start, finish = _c64(frame), _c64(ts), _c64(start), _c64(finish); out = np.empty_like(frame)
    lib().ref_deskew(_d(frame), _d(ts), C.c_size_t(len(frame)), _d(start), _d(finish), _d(out))
    return out


class RefMap:
    """sage_icp::VoxelHashMap, the reference's own (core/VoxelHashMap.hpp)."""

    def __init__(self, voxel_size, max_distance, basic, critical, basic_labels):
        lab = np.ascontiguousarray(basic_labels, dtype=np.int32)
        self.stride = basic + critical
        self.h = C.c_void_p(lib().ref_map_create(C.c_double(voxel_size), C.c_double(max_distance), basic, critical, lab.ctypes.data_as(_ip), len(lab)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_map_destroy(self.h); self.h = None

    def clear(self): lib().ref_map_clear(self.h)
    def empty(self) -> bool: return bool(lib().ref_map_empty(self.h))
    def num_voxels(self) -> int: return int(lib().ref_map_num_voxels(self.h))

    def add_points(self, pts):
        pts = _c64(pts); lib().ref_map_add_points(self.h, _d(pts), C.c_size_t(len(pts)))

    def remove_far(self, origin):
        o = _c64(origin); lib().ref_map_remove_far(self.h, _d(o))

    def update(self, pts, pose):
        pts, pose = _c64(pts), _c64(pose); lib().ref_map_update(self.h, _d(pts), C.c_size_t(len(pts)), _d(pose))

    def pointcloud(self) -> np.ndarray:
        n = lib().ref_map_pointcloud(self.h, None, C.c_size_t(0)); out = np.empty((n, 4))
        lib().ref_map_pointcloud(self.h, _d(out), C.c_size_t(n)); return out

    def dump(self):
        v = self.num_voxels()
        keys = np.zeros((v, 3), np.int32); counts = np.zeros(v, np.int32); pts = np.zeros((v, self.stride, 4))
        lib().ref_map_dump(self.h, keys.ctypes.data_as(_ip), counts.ctypes.data_as(_ip), _d(pts), self.stride, C.c_size_t(v))
        return keys, counts, pts

    def get_correspondences(self, pts, max_dist, th):
        pts = _c64(pts); n = len(pts); src, tgt = np.empty((n, 4)), np.empty((n, 4))
        k = lib().ref_map_get_correspondences(self.h, _d(pts), C.c_size_t(n), C.c_double(max_dist), C.c_double(th), _d(src), _d(tgt))
        return src[:k].copy(), tgt[:k].copy()

    def register_frame_core(self, frame, guess, max_dist, kernel, sem_th) -> np.ndarray:
        frame, guess = _c64(frame), _c64(guess); out = np.empty(7)
        lib().ref_register_frame_core(self.h, _d(frame), C.c_size_t(len(frame)), _d(guess), C.c_double(max_dist), C.c_double(kernel),
                                      C.c_double(sem_th), _d(out))
        return out


class RefPipeline:
    """sage_icp::pipeline::sageICP, the reference's own (pipeline/sageICP.hpp)."""

    def __init__(self, cfg):
        self._pod = cfg.to_pod()
        self.h = C.c_void_p(lib().ref_create(C.byref(self._pod)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_destroy(self.h); self.h = None

    def reset(self): lib().ref_reset(self.h)

    def register_frame(self, pts, timestamps=None) -> np.ndarray:
        pts = _c64(pts); pose = np.empty(7)
        ts = None if timestamps is None else _d(_c64(timestamps))
        lib().ref_register_frame(self.h, _d(pts), C.c_size_t(len(pts)), ts, _d(pose))
        return pose

    def voxelize(self, pts):
        pts = _c64(pts); s, d = np.empty_like(pts), np.empty_like(pts); ns, nd = C.c_size_t(), C.c_size_t()
        lib().ref_voxelize(self.h, _d(pts), C.c_size_t(len(pts)), _d(s), C.byref(ns), _d(d), C.byref(nd))
        return s[:ns.value].copy(), d[:nd.value].copy()

    def _cloud(self, fn):
        n = fn(self.h, None, C.c_size_t(0)); out = np.empty((n, 4))
        fn(self.h, _d(out), C.c_size_t(n)); return out

    def last_source(self): return self._cloud(lib().ref_last_source)
    def local_map(self): return self._cloud(lib().ref_local_map)
    def adaptive_threshold(self) -> float: return float(lib().ref_get_adaptive_threshold(self.h))
    def has_moved(self) -> bool: return bool(lib().ref_has_moved(self.h))

    def prediction_model(self):
        out = np.empty(7); lib().ref_get_prediction_model(self.h, _d(out)); return out

    def poses(self) -> np.ndarray:
        n = int(lib().ref_num_poses(self.h)); out = np.empty((n, 7))
        for i in range(n):
            lib().ref_get_pose(self.h, C.c_size_t(i), _d(out[i]))
        return out

    def transform_to_last_frame(self, last, cur, pts):
        last, cur, pts = _c64(last), _c64(cur), _c64(pts); out = np.empty_like(pts)
        lib().ref_transform_to_last_frame(self.h, _d(last), _d(cur), _d(pts), C.c_size_t(len(pts)), _d(out))
        return out
