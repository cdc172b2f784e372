"""The drop-in boundary without a GPU: the C-ABI library loads, exports every symbol include/sage_icp_b200.h declares,
fails loudly (no CPU fallback) when no sm_100 device is present, and its pure-host entry points work."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import sage_icp_b200 as sg
    sg.build_library()
    return sg.load_library()


def test_header_is_plain_c():
    """No C++ / torch types in the signatures: the header compiles as C99."""
    src = '#include "sage_icp_b200.h"\nint main(void) { return sage_device_count() < 0; }\n'
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c", "-"],
                       input=src.encode(), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()


def test_library_exports_every_declared_symbol(lib):
    from sage_icp_b200.capi import declared_symbols
    syms = declared_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    out = subprocess.run(["nm", "-D", "--defined-only", lib._name], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (sage_[a-z0-9_]+)", out))
    assert set(syms) <= exported
    # the library is self-contained CUDA + C++: it must not pull in torch or the oracle
    needed = subprocess.run(["ldd", lib._name], capture_output=True, text=True).stdout
    assert "torch" not in needed and "oracle" not in needed


def test_sm100a_code_is_in_the_library(lib):
    out = subprocess.run(["cuobjdump", "-lelf", lib._name], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def _has_gpu():
    import sage_icp_b200 as sg
    return sg.device_count() > 0


def test_no_device_means_loud_failure(lib, cfg):
    """The product path has no CPU fallback: on a box without a B200 every create call fails with a message."""
    if _has_gpu():
        pytest.skip("a B200 is visible here")
    import sage_icp_b200 as sg
    assert sg.device_count() == 0
    with pytest.raises(sg.SageError, match="(?i)device|cuda"):
        sg.SageMap(0.8, 100.0, 20, 20, [40])
    with pytest.raises(sg.SageError, match="(?i)device|cuda"):
        sg.SagePipeline(cfg)


def test_package_does_not_import_the_oracle():
    code = "import sys, sage_icp_b200, sage_icp_b200.capi, sage_icp_b200.synthetic; print(any(m.startswith('oracle') for m in sys.modules))"
    out = subprocess.run(["python", "-c", code], capture_output=True, text=True, cwd=ROOT)
    assert out.stdout.strip() == "False", out.stdout + out.stderr
    for dirpath, _, files in os.walk(os.path.join(ROOT, "sage_icp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle_py" not in txt and "liboracle" not in txt and '#include "../../oracle' not in txt, f


def test_shard_range_partitions_exactly(lib):
    import sage_icp_b200 as sg
    for n in (0, 1, 7, 120000, 500000, 2 ** 33 + 5):
        for world in (1, 2, 3, 4, 8):
            edges = [sg.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges, edges[1:]))
            sizes = [e - b for b, e in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(sg.SageError):
        sg.shard_range(10, 2, 2)


def test_transform_to_last_frame_host_entry_point(lib, orc):
    """sageICP::TransformToLastFrame (pipeline/sageICP.cpp:123-129) is pure host arithmetic: callable without a device."""
    rng = np.random.default_rng(0)
    pts = np.c_[rng.normal(0, 10, (50, 3)), rng.choice([40.0, 0.0], 50)]
    a, b = orc.se3_exp(rng.normal(size=6) * 0.3), orc.se3_exp(rng.normal(size=6) * 0.3)
    out = np.empty_like(pts)
    dp = C.POINTER(C.c_double)
    rc = lib.sage_transform_to_last_frame(None, a.ctypes.data_as(dp), b.ctypes.data_as(dp), pts.ctypes.data_as(dp), C.c_size_t(50),
                                          out.ctypes.data_as(dp))
    assert rc == 0
    T = orc.se3_mul(orc.se3_inverse(a), b)
    exp = np.array([orc.se3_act(T, p[:3]) for p in pts])
    assert np.allclose(out[:, :3], exp, atol=1e-12) and np.array_equal(out[:, 3], pts[:, 3])


def test_config_pod_layout_matches_the_header(cfg):
    """ctypes mirror == C struct: compile a probe that prints sizeof/offsetof and compare."""
    from sage_icp_b200.config import ConfigPOD
    fields = [f[0] for f in ConfigPOD._fields_]
    body = "".join(f'printf("{f} %zu\\n", offsetof(sage_config_pod, {f}));' for f in fields)
    src = f'#include <stdio.h>\n#include <stddef.h>\n#include "sage_icp_b200.h"\nint main(void){{ printf("size %zu\\n", sizeof(sage_config_pod)); {body} return 0; }}'
    exe = "/tmp/sage_pod_probe"
    r = subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-x", "c", "-", "-o", exe], input=src.encode(), capture_output=True)
    assert r.returncode == 0, r.stderr.decode()
    out = dict(l.split() for l in subprocess.run([exe], capture_output=True, text=True).stdout.splitlines())
    assert int(out["size"]) == C.sizeof(ConfigPOD)
    for f in fields:
        assert int(out[f]) == getattr(ConfigPOD, f).offset, f
    pod = cfg.to_pod()
    assert pod.n_groups == len(cfg.voxel_labels) == 6 and pod.voxel_size_map == 0.8 and pod.sem_th == 0.4


@pytest.mark.parametrize("n,spread,seed", [(0, 3, 0), (1, 3, 1), (2, 3, 2), (300, 3, 3), (2049, 20, 4), (9000, 60, 5), (40000, 200, 6)])
def test_product_robin_replay_matches_the_oracle_table(lib, orc, n, spread, seed):
    """The product's own host-side replay of the tsl::robin_map order (csrc/frontend.cu, used by VoxelDownsample) against the
    oracle's RobinTable — two independent implementations of the same published rules (and the oracle one is itself pinned
    against a pure-Python emulator in test_oracle_robin.py)."""
    import sage_icp_b200 as sg
    rng = np.random.default_rng(seed)
    keys = np.unique(rng.integers(-spread, spread + 1, size=(4 * max(n, 1), 3)), axis=0)
    rng.shuffle(keys)
    keys = keys[:n].astype(np.int32)
    h = np.array([orc.voxel_hash(*[int(v) for v in k]) for k in keys], dtype=np.uint32)
    got = sg.robin_iteration_order(h)
    want, _ = orc.robin_order(keys) if len(keys) else (np.zeros(0, np.int64), 0)
    assert np.array_equal(got.astype(np.int64), want)


def test_product_robin_replay_with_saturated_hash(lib, orc):
    """Many keys on few hash values: long probe sequences, the regime of SURVEY.md A.9 in miniature."""
    import sage_icp_b200 as sg
    keys = np.array([[(i % 7) + ((i // 7) << 20), 0, 0] for i in range(3000)], dtype=np.int64).astype(np.int32)  # 7 distinct 20-bit hashes
    h = np.array([orc.voxel_hash(int(k[0]), 0, 0) for k in keys], dtype=np.uint32)
    assert len(np.unique(h)) == 7
    got = sg.robin_iteration_order(h)
    want, _ = orc.robin_order(keys)
    assert np.array_equal(got.astype(np.int64), want)


def _colliding_keys(n, hashes, seed):
    """n distinct PACKABLE voxel keys (|coordinate| < 2^20) on the given values of the reference's 20-bit hash: for any (y, z)
    one x mod 2^20 hits a given hash value, because the x multiplier is odd."""
    rng = np.random.default_rng(seed)
    inv = pow(73856093, -1, 1 << 20)
    yz = np.unique(rng.integers(-500, 500, size=(2 * n, 2)), axis=0)
    rng.shuffle(yz)
    keys = []
    for i, (y, z) in enumerate(yz[:n]):
        target = hashes[i % len(hashes)]
        rest = ((int(y) & 0xFFFFFFFF) * 19349663 ^ (int(z) & 0xFFFFFFFF) * 83492791) & 0xFFFFF
        keys.append(((inv * (target ^ rest)) & 0xFFFFF, int(y), int(z)))
    return np.array(keys, dtype=np.int32)


@pytest.mark.parametrize("n,spread,seed", [(1, 3, 0), (40, 3, 1), (700, 8, 2), (5000, 25, 3)])
def test_host_mirror_table_replay_matches_python_model(lib, n, spread, seed):
    """The host mirror of the map's tsl::robin_map (faithful-eviction mode, csrc/voxel_map.cu HostVoxelTable): inserts, the
    erase-while-iterating sweep and clear() against the pure-Python model of tests/test_oracle_robin.py, op for op."""
    import sage_icp_b200 as sg
    from test_oracle_robin import PyRobinMap, _hash
    rng = np.random.default_rng(seed)
    t, live, ops = PyRobinMap(), set(), []
    origin = np.zeros(3, dtype=np.int64)
    for rnd in range(6):
        fresh = np.unique(rng.integers(-spread, spread + 1, size=(n, 3)) + origin, axis=0)
        rng.shuffle(fresh)
        for k in fresh:
            k = tuple(int(v) for v in k)
            if k in live:
                continue
            live.add(k)
            t.insert(_hash(k), k)
            ops.append((0, *k, 0))
        r2 = int((0.8 * spread) ** 2) + 1
        far = lambda k, o=origin.copy(): sum((a - int(b)) ** 2 for a, b in zip(k, o)) > r2
        t.sweep(far)
        ops.append((1, *[int(v) for v in origin], r2))
        live = set(t.order())
        if rnd == 3:
            t.b = [None] * len(t.b)  # clear(): buckets emptied, bucket count kept
            live = set()
            ops.append((2, 0, 0, 0, 0))
        keys, buckets = sg.robin_table_replay(ops)
        assert [tuple(int(v) for v in k) for k in keys] == t.order(), rnd
        assert buckets == len(t.b)
        origin += rng.integers(0, spread // 2 + 2, 3)
    assert len(sg.robin_table_replay(np.zeros((0, 5)))[0]) == 0
    with pytest.raises(sg.SageError):
        sg.robin_table_replay([(7, 0, 0, 0, 0)])


@pytest.mark.parametrize("n,hashes", [(3000, list(range(7))), (8300, [77 + (j << 15) for j in range(32)])])
def test_host_mirror_table_with_saturated_hash(lib, orc, n, hashes):
    """Long probe sequences.  Second case: 8300 keys whose hashes agree in the low 15 bits collide on ONE bucket while the table
    has <= 2^15 buckets, so the probe-length limit (8192) forces a growth the load factor alone would not ask for; compared with
    the oracle's RobinTable (iteration order and bucket count).  (Keys on a single hash value would double forever — in tsl too.)"""
    import sage_icp_b200 as sg
    keys = _colliding_keys(n, hashes, 5)
    assert set(orc.voxel_hash(*[int(v) for v in k]) for k in keys[:200]) <= set(hashes)
    got, buckets = sg.robin_table_replay(np.c_[np.zeros(len(keys), np.int32), keys, np.zeros(len(keys), np.int32)])
    want, want_buckets = orc.robin_order(keys)
    assert np.array_equal(got, keys[want])
    assert buckets == want_buckets
    if len(hashes) == 32:
        assert buckets == 2 * (1 << int(np.ceil(np.log2(2 * n - 1))))  # one doubling more than the load factor alone


def test_null_handles_are_error_codes_not_crashes(lib):
    """Every entry point checks its handle: a NULL pipeline / map gives a negative code and a message."""
    L = lib
    d7 = (C.c_double * 7)()
    null = C.c_void_p(None)
    for name, args in [("sage_reset", ()), ("sage_last_iterations", ()), ("sage_has_moved", ()), ("sage_get_prediction_model", (d7,)),
                       ("sage_get_pose", (C.c_size_t(0), d7)), ("sage_set_devices", ((C.c_int * 1)(0), 1)), ("sage_register_frame", (None, C.c_size_t(0), None, d7, None, None))]:
        rc = getattr(L, name)(null, *args)
        assert rc < 0, name
        assert b"null" in L.sage_last_error(), name
    for name, args in [("sage_map_clear", ()), ("sage_map_empty", ()), ("sage_map_set_eviction", (1,)), ("sage_map_remove_far", (d7,)),
                       ("sage_map_add_points", (None, C.c_size_t(0))), ("sage_map_comm_destroy", ()), ("sage_map_profile_enable", (1,))]:
        rc = getattr(L, name)(null, *args)
        assert rc < 0, name
        assert b"null" in L.sage_last_error(), name
    L.sage_num_poses.restype = C.c_int64
    L.sage_local_map.restype = C.c_int64
    assert L.sage_num_poses(null) < 0 and L.sage_local_map(null, None, C.c_size_t(0)) < 0
    L.sage_pipeline_map.restype = C.c_void_p
    L.sage_map_stream.restype = C.c_void_p
    assert L.sage_pipeline_map(null) is None and L.sage_map_stream(null) is None
    L.sage_destroy(null); L.sage_map_destroy(null)  # no-ops


def test_product_robin_replay_at_the_probe_length_limit(lib, orc):
    """The down-sampler's replay (csrc/frontend.cu) where the probe-length limit (8192) forces a growth: 8300 keys whose hashes
    agree in the low 15 bits; same order as the oracle's RobinTable."""
    import sage_icp_b200 as sg
    keys = _colliding_keys(8300, [77 + (j << 15) for j in range(32)], 6)
    h = np.array([orc.voxel_hash(*[int(v) for v in k]) for k in keys], dtype=np.uint32)
    got = sg.robin_iteration_order(h)
    want, buckets = orc.robin_order(keys)
    assert buckets == 1 << 16
    assert np.array_equal(got.astype(np.int64), want)


def test_robin_tables_agree_when_the_new_element_itself_lands_at_the_limit(lib, orc):
    """Corner of the growth rule: the inserted element's first swap happens at distance exactly 8192 (a second cluster sits right
    behind an 8192-long run).  tsl only marks the table for growth when a CARRIED element is swapped at >= 8192, so no extra
    growth here — the oracle's table, the down-sampler's replay and the map's host mirror must all agree."""
    import sage_icp_b200 as sg
    c = 77
    run = _colliding_keys(8300, [c + (j << 15) for j in range(32)], 7)        # ideal bucket c while the table has 2^15 buckets
    behind = _colliding_keys(40, [c + 8192 + (j << 15) for j in range(32)], 8)  # ideal bucket c + 8192: right behind the run
    after = _colliding_keys(10, [5], 9)  # inserted after the corner case: a table wrongly marked for growth would double now
    keys = np.concatenate([run[:4000], behind, run[4000:8193], after])
    assert len(np.unique(keys, axis=0)) == len(keys)
    h = np.array([orc.voxel_hash(*[int(v) for v in k]) for k in keys], dtype=np.uint32)
    want, buckets = orc.robin_order(keys)
    got = sg.robin_iteration_order(h)
    assert np.array_equal(got.astype(np.int64), want)
    got2, buckets2 = sg.robin_table_replay(np.c_[np.zeros(len(keys), np.int32), keys, np.zeros(len(keys), np.int32)])
    assert buckets2 == buckets == 1 << 15 and np.array_equal(got2, keys[want])


def test_bad_config_is_refused_before_any_device_work(lib, cfg):
    """sage_create validates the POD first (NULL arrays, negative lengths): a message, never a crash — also on a box without a GPU."""
    L = lib
    L.sage_create.restype = C.c_void_p
    assert L.sage_create(None, 0) is None and b"NULL" in L.sage_last_error()
    pod = cfg.to_pod()
    keep = pod.group_labels
    pod.group_labels = None
    assert L.sage_create(C.byref(pod), 0) is None and b"group_labels" in L.sage_last_error()
    pod.group_labels = keep
    pod.n_basic_parts_labels = -1
    assert L.sage_create(C.byref(pod), 0) is None and b"negative" in L.sage_last_error()
    pod = cfg.to_pod()
    keep = pod.voxel_size
    pod.voxel_size = None
    assert L.sage_create(C.byref(pod), 0) is None and b"voxel_size" in L.sage_last_error()


def test_python_harness_refuses_inconsistent_inputs():
    """The ctypes harness must not hand the library lengths it has not checked: a group count taken from voxel_size with offsets taken
    from voxel_labels, or a flat array passed as a cloud, would make the library read past the end of a buffer."""
    from sage_icp_b200.capi import _pts
    from sage_icp_b200.config import launch_config
    cfg = launch_config()
    cfg.voxel_size = cfg.voxel_size + [0.5]  # one more size than label groups
    with pytest.raises(ValueError, match="voxel_labels has 6 groups but voxel_size has 7"):
        cfg.to_pod()
    with pytest.raises(ValueError, match=r"\(n, 4\)"):
        _pts(np.zeros(12))
    with pytest.raises(ValueError, match=r"\(n, 4\)"):
        _pts(np.zeros((5, 3)))
    assert _pts(np.zeros((0,))).shape == (0, 4) and _pts([[1, 2, 3, 4]]).shape == (1, 4)
