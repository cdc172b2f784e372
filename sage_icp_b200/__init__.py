"""sage_icp_b200 — B200-native SAGE-ICP registration hot path.

The product is the C-ABI shared library ``lib/libsage_icp_b200.so`` (hand-written sm_100a CUDA + C++ host code,
sources in ``csrc/``; header ``include/sage_icp_b200.h``).  This Python package is a thin ctypes harness over that
ABI used by the tests and ``bench.py``; the C++ drop-in for the ROS node is ``include/sage_icp/pipeline/sageICP.hpp``.
There is no CPU fallback: every entry point raises if the CUDA library or a B200 is missing.
"""
from .config import SageConfig, launch_config  # noqa: F401
from .capi import (  # noqa: F401
    SagePipeline, SageMap, SageError, build_library, library_path, load_library, device_count, launch_count,
    shard_range, nccl_unique_id, robin_iteration_order, robin_table_replay,
)
