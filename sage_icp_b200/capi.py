"""ctypes binding of include/sage_icp_b200.h (the C ABI the reference's C++ adaptor binds too).

Class/method names mirror the reference: ``SagePipeline`` ~ sage_icp::pipeline::sageICP
(pipeline/sageICP.hpp:67-109), ``SageMap`` ~ sage_icp::VoxelHashMap (core/VoxelHashMap.hpp:34-106) plus the free
function sage_icp::RegisterFrame (core/Registration.hpp:33-38) as ``SageMap.register_frame``.
"""
from __future__ import annotations

import ctypes as C
import os
import re
import subprocess
from typing import List, Optional, Tuple

import numpy as np

from .config import ConfigPOD, SageConfig

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.environ.get("SAGE_ICP_LIB") or os.path.join(_HERE, "lib", "libsage_icp_b200.so")  # SAGE_ICP_LIB: tuning builds
_HEADER = os.path.join(os.path.dirname(_HERE), "include", "sage_icp_b200.h")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_u8p = C.POINTER(C.c_uint8)


class SageError(RuntimeError):
    pass


def library_path() -> str:
    return _LIB


def build_library(force: bool = False) -> str:
    """Compile csrc/*.cu for sm_100a into lib/libsage_icp_b200.so (nvcc cross-compiles without a GPU)."""
    src = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(src, f) for f in os.listdir(src)] + [_HEADER]
    stale = (not os.path.exists(_LIB)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", src, "-j8"] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return _LIB


def declared_symbols() -> List[str]:
    """Every function name include/sage_icp_b200.h declares."""
    txt = open(_HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(sage_[a-z0-9_]+)\s*\(", txt)))


_lib = None


def load_library() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB):
        raise SageError(f"{_LIB} is missing: build it with __graft_entry__.build() / make -C sage_icp_b200/csrc "
                        "(there is no CPU fallback)")
    L = C.CDLL(_LIB)
    L.sage_last_error.restype = C.c_char_p
    L.sage_create.restype = C.c_void_p
    L.sage_map_create.restype = C.c_void_p
    L.sage_pipeline_map.restype = C.c_void_p
    L.sage_map_stream.restype = C.c_void_p
    for f in ("sage_last_source", "sage_last_frame_downsample", "sage_num_poses", "sage_get_poses", "sage_local_map", "sage_map_num_voxels",
              "sage_map_num_points", "sage_map_pointcloud", "sage_map_dump", "sage_map_get_correspondences", "sage_preprocess",
              "sage_voxel_downsample", "sage_launch_count"):
        getattr(L, f).restype = C.c_int64
    for f in ("sage_last_sigma", "sage_get_adaptive_threshold"):
        getattr(L, f).restype = C.c_double
    _lib = L
    return L


def _err(L, what: str, rc) -> SageError:
    msg = L.sage_last_error()
    return SageError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}")


def _d(a: np.ndarray):
    return a.ctypes.data_as(_dp)


def _c64(a) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a


def _pts(a) -> np.ndarray:
    """A cloud as the C ABI takes it: contiguous f64, n x 4 (x, y, z, label).  len() of the result is the point count handed to
    the library, so anything that is not n x 4 is refused here rather than over-read there."""
    a = _c64(a)
    if a.ndim == 1 and a.size == 0:
        return a.reshape(0, 4)
    if a.ndim != 2 or a.shape[1] != 4:
        raise ValueError(f"expected an (n, 4) array of x, y, z, label; got shape {a.shape}")
    return a


def device_count() -> int:
    return int(load_library().sage_device_count())


def launch_count() -> int:
    return int(load_library().sage_launch_count())


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    L = load_library()
    b, e = C.c_size_t(), C.c_size_t()
    rc = L.sage_shard_range(C.c_size_t(n), rank, world, C.byref(b), C.byref(e))
    if rc != 0:
        raise _err(L, "sage_shard_range", rc)
    return int(b.value), int(e.value)


def robin_iteration_order(hash20) -> np.ndarray:
    """Host utility of the down-sampler: tsl::robin_map iteration order for distinct keys with these 20-bit hashes."""
    L = load_library()
    h = np.ascontiguousarray(hash20, dtype=np.uint32); out = np.empty(len(h), dtype=np.uint32)
    rc = L.sage_robin_iteration_order(h.ctypes.data_as(C.POINTER(C.c_uint32)), C.c_size_t(len(h)), out.ctypes.data_as(C.POINTER(C.c_uint32)))
    if rc != 0:
        raise _err(L, "sage_robin_iteration_order", rc)
    return out


def robin_table_replay(ops):
    """Host mirror table of the faithful-eviction mode (no device): ops = rows of (op, x, y, z, r2); returns (keys in
    iteration order, bucket count)."""
    L = load_library()
    L.sage_robin_table_replay.restype = C.c_int64
    o = np.ascontiguousarray(ops, dtype=np.int32).reshape(-1, 5)
    out = np.empty((len(o), 3), dtype=np.int32); bc = C.c_uint64(0)
    ip = C.POINTER(C.c_int32)
    k = L.sage_robin_table_replay(o.ctypes.data_as(ip), C.c_size_t(len(o)), out.ctypes.data_as(ip), C.c_size_t(len(o)), C.byref(bc))
    if k < 0:
        raise _err(L, "sage_robin_table_replay", int(k))
    return out[:k].copy(), int(bc.value)


def nccl_unique_id() -> bytes:
    L = load_library()
    buf = (C.c_uint8 * 128)()
    rc = L.sage_nccl_unique_id(buf)
    if rc != 0:
        raise _err(L, "sage_nccl_unique_id", rc)
    return bytes(buf)


class SageMap:
    """sage_icp::VoxelHashMap on the device."""

    def __init__(self, voxel_size: float, max_distance: float, basic: int, critical: int, basic_labels, device: int = 0,
                 _borrow: Optional[int] = None, _keepalive=None):
        self.L = load_library()
        self.stride = basic + critical
        self._own = _borrow is None
        self._keepalive = _keepalive
        if _borrow is None:
            lab = (C.c_int32 * max(1, len(basic_labels)))(*basic_labels)
            h = self.L.sage_map_create(C.c_double(voxel_size), C.c_double(max_distance), basic, critical, lab, len(basic_labels), device)
            if not h:
                raise _err(self.L, "sage_map_create", None)
            self.h = C.c_void_p(h)
        else:
            self.h = C.c_void_p(_borrow)

    def __del__(self):
        if getattr(self, "_own", False) and getattr(self, "h", None):
            self.L.sage_map_destroy(self.h)
            self.h = None

    def _chk(self, rc, what):
        if rc < 0:
            raise _err(self.L, what, rc)
        return rc

    def clear(self): self._chk(self.L.sage_map_clear(self.h), "sage_map_clear")

    def set_eviction(self, faithful: bool):
        """True: the reference's erase-while-iterating eviction and robin_map iteration order (empty map only)."""
        self._chk(self.L.sage_map_set_eviction(self.h, int(bool(faithful))), "sage_map_set_eviction")
    def empty(self) -> bool: return bool(self._chk(self.L.sage_map_empty(self.h), "sage_map_empty"))
    def num_voxels(self) -> int: return int(self._chk(self.L.sage_map_num_voxels(self.h), "sage_map_num_voxels"))
    def num_points(self) -> int: return int(self._chk(self.L.sage_map_num_points(self.h), "sage_map_num_points"))

    def add_points(self, pts):
        pts = _pts(pts)
        self._chk(self.L.sage_map_add_points(self.h, _d(pts), C.c_size_t(len(pts))), "sage_map_add_points")

    def remove_far(self, origin):
        o = _c64(origin)
        self._chk(self.L.sage_map_remove_far(self.h, _d(o)), "sage_map_remove_far")

    def update(self, pts, pose):
        pts, pose = _pts(pts), _c64(pose)
        self._chk(self.L.sage_map_update(self.h, _d(pts), C.c_size_t(len(pts)), _d(pose)), "sage_map_update")

    def pointcloud(self) -> np.ndarray:
        n = self.num_points()
        out = np.empty((n, 4))
        k = self._chk(self.L.sage_map_pointcloud(self.h, _d(out), C.c_size_t(n)), "sage_map_pointcloud")
        return out[:k]

    def load(self, keys, counts, pts):
        keys = np.ascontiguousarray(keys, np.int32); counts = np.ascontiguousarray(counts, np.int32); pts = _c64(pts)
        stride = pts.shape[1] if pts.ndim == 3 else self.stride
        self._chk(self.L.sage_map_load(self.h, keys.ctypes.data_as(_ip), counts.ctypes.data_as(_ip), _d(pts), int(stride),
                                       C.c_size_t(len(counts))), "sage_map_load")

    def dump(self):
        v = self.num_voxels()
        keys = np.zeros((v, 3), np.int32); counts = np.zeros(v, np.int32); pts = np.zeros((v, self.stride, 4))
        k = self._chk(self.L.sage_map_dump(self.h, keys.ctypes.data_as(_ip), counts.ctypes.data_as(_ip), _d(pts), C.c_size_t(v)),
                      "sage_map_dump")
        return keys[:k], counts[:k], pts[:k]

    def get_correspondences(self, pts, max_dist: float, th: float):
        """Per-query result: (target (n,4), matched (n,) bool)."""
        pts = _pts(pts); n = len(pts)
        tgt = np.zeros((n, 4)); matched = np.zeros(n, np.uint8)
        self._chk(self.L.sage_map_get_correspondences(self.h, _d(pts), C.c_size_t(n), C.c_double(max_dist), C.c_double(th), _d(tgt),
                                                      matched.ctypes.data_as(_u8p)), "sage_map_get_correspondences")
        return tgt, matched.astype(bool)

    def nn_stats(self, pts) -> Tuple[int, int]:
        pts = _pts(pts); o, c = C.c_uint64(), C.c_uint64()
        self._chk(self.L.sage_map_nn_stats(self.h, _d(pts), C.c_size_t(len(pts)), C.byref(o), C.byref(c)), "sage_map_nn_stats")
        return int(o.value), int(c.value)

    def search_work(self, pts, max_dist: float, th: float, with_staged: bool = False):
        """(records ranked, table probes, queries re-ranked in f64, queries finished by a whole warp[, records staged through
        TMA bulk copies — tile search only]) of one correspondence pass."""
        pts = _pts(pts); a, b, c, d, e = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
        self._chk(self.L.sage_map_search_work(self.h, _d(pts), C.c_size_t(len(pts)), C.c_double(max_dist), C.c_double(th),
                                              C.byref(a), C.byref(b), C.byref(c), C.byref(d), C.byref(e)), "sage_map_search_work")
        out = (int(a.value), int(b.value), int(c.value), int(d.value))
        return out + (int(e.value),) if with_staged else out

    def normal_equations(self, pts, max_dist: float, kernel: float, sem_th: float):
        pts = _pts(pts); JTJ = np.zeros((6, 6)); JTr = np.zeros(6); n = C.c_int64()
        self._chk(self.L.sage_core_normal_equations(self.h, _d(pts), C.c_size_t(len(pts)), C.c_double(max_dist), C.c_double(kernel),
                                                    C.c_double(sem_th), _d(JTJ), _d(JTr), C.byref(n)), "sage_core_normal_equations")
        return JTJ, JTr, int(n.value)

    def register_frame(self, frame, guess, max_dist: float, kernel: float, sem_th: float, max_iters: int = 0, est_th: float = -1.0):
        """sage_icp::RegisterFrame (core/Registration.cpp:113-141) with HOST buffers.  Returns (pose7, iterations)."""
        frame, guess = _pts(frame), _c64(guess); out = np.empty(7); it = C.c_int()
        self._chk(self.L.sage_core_register_frame(self.h, _d(frame), C.c_size_t(len(frame)), _d(guess), C.c_double(max_dist),
                                                  C.c_double(kernel), C.c_double(sem_th), max_iters, C.c_double(est_th), _d(out),
                                                  C.byref(it)), "sage_core_register_frame")
        return out, int(it.value)

    def register_frame_device(self, dev_ptr: int, n: int, guess, max_dist: float, kernel: float, sem_th: float, max_iters: int = 0,
                              est_th: float = -1.0):
        """Same with a DEVICE-resident frame (n x 4 f64), e.g. ``torch_tensor.data_ptr()``."""
        guess = _c64(guess); out = np.empty(7); it = C.c_int()
        self._chk(self.L.sage_core_register_frame_device(self.h, C.c_void_p(dev_ptr), C.c_size_t(n), _d(guess), C.c_double(max_dist),
                                                         C.c_double(kernel), C.c_double(sem_th), max_iters, C.c_double(est_th),
                                                         _d(out), C.byref(it)), "sage_core_register_frame_device")
        return out, int(it.value)

    def stream(self) -> int:
        return int(self.L.sage_map_stream(self.h) or 0)

    def profile_enable(self, on: bool): self.L.sage_map_profile_enable(self.h, int(on))

    def profile_read(self) -> Tuple[int, float]:
        """(Gauss-Newton iterations timed, their total device milliseconds) since the last read."""
        n, ms = C.c_int64(), C.c_double()
        self._chk(self.L.sage_map_profile_read(self.h, C.byref(n), C.byref(ms)), "sage_map_profile_read")
        return int(n.value), float(ms.value)

    def profile_read_launches(self) -> Tuple[int, int, float]:
        """(iterations, kernel launches, total device milliseconds): a cooperative launch runs a whole registration's loop."""
        n, k, ms = C.c_int64(), C.c_int64(), C.c_double()
        self._chk(self.L.sage_map_profile_read_launches(self.h, C.byref(n), C.byref(k), C.byref(ms)), "sage_map_profile_read_launches")
        return int(n.value), int(k.value), float(ms.value)

    def comm_init(self, rank: int, world: int, uid: bytes):
        buf = (C.c_uint8 * 128).from_buffer_copy(uid)
        self._chk(self.L.sage_map_comm_init(self.h, rank, world, buf), "sage_map_comm_init")

    def comm_peer_handle(self) -> bytes:
        buf = (C.c_uint8 * 64)()
        self._chk(self.L.sage_map_comm_peer_handle(self.h, buf), "sage_map_comm_peer_handle")
        return bytes(buf)

    def comm_peer_attach(self, rank: int, world: int, handles: bytes):
        assert len(handles) == 64 * world
        buf = (C.c_uint8 * len(handles)).from_buffer_copy(handles)
        self._chk(self.L.sage_map_comm_peer_attach(self.h, rank, world, buf), "sage_map_comm_peer_attach")

    def comm_destroy(self): self._chk(self.L.sage_map_comm_destroy(self.h), "sage_map_comm_destroy")


class SagePipeline:
    """sage_icp::pipeline::sageICP on the device."""

    def __init__(self, cfg: SageConfig, device: int = 0):
        self.L = load_library()
        self.cfg = cfg
        self._pod: ConfigPOD = cfg.to_pod()
        h = self.L.sage_create(C.byref(self._pod), device)
        if not h:
            raise _err(self.L, "sage_create", None)
        self.h = C.c_void_p(h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.sage_destroy(self.h)
            self.h = None

    def _chk(self, rc, what):
        if rc < 0:
            raise _err(self.L, what, rc)
        return rc

    def reinitialize(self): self._chk(self.L.sage_reset(self.h), "sage_reset")

    def set_devices(self, ids):
        """A fresh pipeline on GPU ids[0]; with several ids, one replica of the map per GPU and every frame's ICP queries sharded
        over all of them (one process, peer-memory all-reduce inside the search kernel)."""
        a = (C.c_int * len(ids))(*ids)
        self._chk(self.L.sage_set_devices(self.h, a, len(ids)), "sage_set_devices")

    def num_devices(self) -> int:
        return int(self._chk(self.L.sage_num_devices(self.h), "sage_num_devices"))

    def register_frame(self, pts, timestamps=None):
        """RegisterFrame(frame[, timestamps]) -> (pose7, t_icp, t_all); `source` via last_source()."""
        pts = _pts(pts); pose = np.empty(7); ti, ta = C.c_double(), C.c_double()
        ts = None if timestamps is None else _d(_c64(timestamps))
        self._chk(self.L.sage_register_frame(self.h, _d(pts), C.c_size_t(len(pts)), ts, _d(pose), C.byref(ti), C.byref(ta)),
                  "sage_register_frame")
        return pose, ti.value, ta.value

    def register_frame_pointcloud2(self, data: np.ndarray, n_points: int, point_step: int, offsets=(0, 4, 8, 12),
                                   label_datatype: int = 2, timestamps=None):
        """RegisterFrame on the raw sensor_msgs/PointCloud2 data buffer (uint8 array), unpacked on the device."""
        data = np.ascontiguousarray(data, dtype=np.uint8); pose = np.empty(7); ti, ta = C.c_double(), C.c_double()
        assert data.size >= n_points * point_step
        ts = None if timestamps is None else _d(_c64(timestamps))
        self._chk(self.L.sage_register_frame_pointcloud2(self.h, data.ctypes.data_as(_u8p), C.c_size_t(n_points), C.c_uint32(point_step),
                                                         C.c_uint32(offsets[0]), C.c_uint32(offsets[1]), C.c_uint32(offsets[2]),
                                                         C.c_uint32(offsets[3]), int(label_datatype), ts, _d(pose), C.byref(ti), C.byref(ta)),
                  "sage_register_frame_pointcloud2")
        return pose, ti.value, ta.value

    def _cloud(self, fn, what):
        n = self._chk(fn(self.h, None, C.c_size_t(0)), what)
        out = np.empty((n, 4))
        k = self._chk(fn(self.h, _d(out), C.c_size_t(n)), what)
        return out[:k]

    def last_source(self): return self._cloud(self.L.sage_last_source, "sage_last_source")
    def last_frame_downsample(self): return self._cloud(self.L.sage_last_frame_downsample, "sage_last_frame_downsample")
    def local_map(self): return self._cloud(self.L.sage_local_map, "sage_local_map")
    def last_iterations(self) -> int: return int(self.L.sage_last_iterations(self.h))
    def last_sigma(self) -> float: return float(self.L.sage_last_sigma(self.h))
    def adaptive_threshold(self) -> float: return float(self.L.sage_get_adaptive_threshold(self.h))
    def has_moved(self) -> bool: return bool(self.L.sage_has_moved(self.h))

    def prediction_model(self):
        out = np.empty(7); self.L.sage_get_prediction_model(self.h, _d(out)); return out

    def voxelize(self, pts):
        pts = _pts(pts); s, d = np.empty_like(pts), np.empty_like(pts); ns, nd = C.c_size_t(), C.c_size_t()
        self._chk(self.L.sage_voxelize(self.h, _d(pts), C.c_size_t(len(pts)), _d(s), C.byref(ns), _d(d), C.byref(nd)), "sage_voxelize")
        return s[:ns.value].copy(), d[:nd.value].copy()

    def preprocess(self, pts):
        pts = _pts(pts); out = np.empty_like(pts)
        n = self._chk(self.L.sage_preprocess(self.h, _d(pts), C.c_size_t(len(pts)), _d(out), C.c_size_t(len(pts))), "sage_preprocess")
        return out[:n].copy()

    def voxel_downsample(self, pts, vox_scale: float):
        pts = _pts(pts); out = np.empty_like(pts)
        n = self._chk(self.L.sage_voxel_downsample(self.h, _d(pts), C.c_size_t(len(pts)), C.c_double(vox_scale), _d(out),
                                                   C.c_size_t(len(pts))), "sage_voxel_downsample")
        return out[:n].copy()

    def transform_to_last_frame(self, last_pose, current_pose, pts):
        last_pose, current_pose, pts = _c64(last_pose), _c64(current_pose), _pts(pts); out = np.empty_like(pts)
        self.L.sage_transform_to_last_frame(self.h, _d(last_pose), _d(current_pose), _d(pts), C.c_size_t(len(pts)), _d(out))
        return out

    def key_frame_grid(self, pts, bounds, rows: int, cols: int, last_pose=None, current_pose=None, last_occ=None):
        """utils::EigenToGridMap (+ TransformToLastFrame first, + compute_occ_overlap against last_occ) on the device.
        Returns (grid[rows, cols] int32, overlap or None)."""
        pts = _pts(pts); b = _c64(np.asarray(bounds, float).reshape(6)); grid = np.zeros((rows, cols), np.int32); ov = C.c_double()
        lp = None if last_pose is None else _d(_c64(last_pose))
        cp = None if current_pose is None else _d(_c64(current_pose))
        lo = None if last_occ is None else np.ascontiguousarray(last_occ, np.int32)
        self._chk(self.L.sage_key_frame_grid(self.h, _d(pts), C.c_size_t(len(pts)), lp, cp, _d(b), rows, cols,
                                             None if lo is None else lo.ctypes.data_as(_ip), grid.ctypes.data_as(_ip), C.byref(ov)),
                  "sage_key_frame_grid")
        return grid, (float(ov.value) if lo is not None else None)

    def poses(self, first: int = 0) -> np.ndarray:
        """poses()[first:] in one bulk call (sage_get_poses)."""
        n = self._chk(self.L.sage_get_poses(self.h, C.c_size_t(first), None, C.c_size_t(0)), "sage_get_poses")
        out = np.empty((n, 7))
        if n:
            self._chk(self.L.sage_get_poses(self.h, C.c_size_t(first), _d(out), C.c_size_t(n)), "sage_get_poses")
        return out

    def pose(self, i: int) -> np.ndarray:
        out = np.empty(7)
        self._chk(self.L.sage_get_pose(self.h, C.c_size_t(i), _d(out)), "sage_get_pose")
        return out

    def map(self) -> SageMap:
        c = self.cfg
        return SageMap(c.voxel_size_map, c.local_map_range, c.basic_points_per_voxel, c.critical_points_per_voxel,
                       c.basic_parts_labels, _borrow=self.L.sage_pipeline_map(self.h), _keepalive=self)
