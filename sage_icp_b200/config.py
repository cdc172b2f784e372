"""sageConfig mirror + its C POD (include/sage_icp_b200.h: sage_config_pod).

Field names, types and defaults follow the reference's ``sage_icp::pipeline::sageConfig``
(cpp/sage_icp/pipeline/sageICP.hpp:39-65).  ``launch_config()`` returns the parameter set the reference's
ROS launch file actually runs with (ros/launch/odometry.launch.py:32-67) with the dynamic-vehicle filter off
(odometry_gt.launch.py:50), which is what every BASELINE.json config uses.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import List


class ConfigPOD(C.Structure):
    """Layout of ``sage_config_pod`` (and, deliberately, of the oracle's ``orc_config_pod``)."""

    _fields_ = [
        ("n_groups", C.c_int32),
        ("group_offsets", C.POINTER(C.c_int32)),
        ("group_labels", C.POINTER(C.c_int32)),
        ("voxel_size", C.POINTER(C.c_double)),
        ("voxel_size_map", C.c_double),
        ("max_range", C.c_double),
        ("min_range", C.c_double),
        ("label_max_range", C.c_double),
        ("local_map_range", C.c_double),
        ("basic_points_per_voxel", C.c_int32),
        ("critical_points_per_voxel", C.c_int32),
        ("n_basic_parts_labels", C.c_int32),
        ("basic_parts_labels", C.POINTER(C.c_int32)),
        ("min_motion_th", C.c_double),
        ("initial_threshold", C.c_double),
        ("sem_th", C.c_double),
        ("deskew", C.c_int32),
        ("dynamic_vehicle_filter", C.c_int32),
        ("dynamic_vehicle_filter_th", C.c_double),
        ("dynamic_vehicle_voxid", C.c_int32),
        ("n_dynamic_remove_lankmark", C.c_int32),
        ("dynamic_remove_lankmark", C.POINTER(C.c_int32)),
    ]


@dataclass
class SageConfig:
    voxel_labels: List[List[int]] = field(default_factory=list)
    voxel_size: List[float] = field(default_factory=list)
    voxel_size_map: float = 1.0
    max_range: float = 100.0
    min_range: float = 5.0
    label_max_range: float = 50.0
    local_map_range: float = 100.0
    basic_points_per_voxel: int = 20
    critical_points_per_voxel: int = 20
    basic_parts_labels: List[int] = field(default_factory=list)
    min_motion_th: float = 0.1
    initial_threshold: float = 2.0
    sem_th: float = 0.4
    deskew: bool = False
    dynamic_vehicle_filter: bool = False
    dynamic_vehicle_filter_th: float = 0.5
    dynamic_vehicle_voxid: int = 5
    dynamic_remove_lankmark: List[int] = field(default_factory=list)

    def to_pod(self) -> ConfigPOD:
        """Build the POD; the backing arrays are kept alive on the returned object (``_keep``)."""
        if len(self.voxel_labels) != len(self.voxel_size):
            # n_groups is taken from voxel_size and the offsets from voxel_labels: a mismatch would make the library read past
            # the offsets array (the C++ adaptor rejects it the same way)
            raise ValueError(f"voxel_labels has {len(self.voxel_labels)} groups but voxel_size has {len(self.voxel_size)} entries")
        offs = [0]
        flat: List[int] = []
        for g in self.voxel_labels:
            flat.extend(int(v) for v in g)
            offs.append(len(flat))
        a_offs = (C.c_int32 * len(offs))(*offs)
        a_flat = (C.c_int32 * max(1, len(flat)))(*flat)
        a_vs = (C.c_double * max(1, len(self.voxel_size)))(*self.voxel_size)
        a_bp = (C.c_int32 * max(1, len(self.basic_parts_labels)))(*self.basic_parts_labels)
        a_lm = (C.c_int32 * max(1, len(self.dynamic_remove_lankmark)))(*self.dynamic_remove_lankmark)
        pod = ConfigPOD(
            len(self.voxel_size), a_offs, a_flat, a_vs,
            self.voxel_size_map, self.max_range, self.min_range, self.label_max_range, self.local_map_range,
            self.basic_points_per_voxel, self.critical_points_per_voxel,
            len(self.basic_parts_labels), a_bp,
            self.min_motion_th, self.initial_threshold, self.sem_th,
            int(self.deskew), int(self.dynamic_vehicle_filter), self.dynamic_vehicle_filter_th,
            self.dynamic_vehicle_voxid, len(self.dynamic_remove_lankmark), a_lm,
        )
        pod._keep = (a_offs, a_flat, a_vs, a_bp, a_lm)
        return pod


def launch_config(**overrides) -> SageConfig:
    """ros/launch/odometry.launch.py:32-67 with dynamic_vehicle_filter=False (odometry_gt.launch.py:50)."""
    cfg = SageConfig(
        voxel_labels=[[40, 44, 48, 49], [50, 51, 52], [70, 72], [60, 71, 80, 81, 99], [0],
                      [10, 11, 13, 15, 16, 18, 20]],
        voxel_size=[0.6, 1.0, 0.9, 0.8, 1.0, 0.6],
        voxel_size_map=0.8, max_range=100.0, min_range=5.0, label_max_range=50.0, local_map_range=100.0,
        basic_points_per_voxel=20, critical_points_per_voxel=20,
        basic_parts_labels=[40, 44, 48, 49, 50, 70, 72],
        min_motion_th=0.1, initial_threshold=2.0, sem_th=0.4,
        deskew=False, dynamic_vehicle_filter=False, dynamic_vehicle_filter_th=0.5, dynamic_vehicle_voxid=5,
        dynamic_remove_lankmark=[44, 48],
    )
    for k, v in overrides.items():
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)
    return cfg
