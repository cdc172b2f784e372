"""bench.py's driver contract, as far as it can be checked without a GPU: the reference arm runs on CPU and prints exactly one
JSON line with the required keys; the sharding helper partitions a scan."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--map-points", "60000", "--cpu-fraction", "0.02"], capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "scans/s" and d["value"] > 0 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"]
    # the value is the faster of the oracle port and the reference's own code (oracle/_ref), and the line says which
    both = [v for v in (cb["port_value"], cb["reference_build_value"]) if v is not None]
    assert d["value"] == pytest.approx(max(both))
    assert (cb["kind"] == "reference") == (cb["reference_build_value"] is not None and cb["reference_build_value"] >= cb["port_value"])
    assert d["e2e"] == {"value": d["value"], "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, cwd=ROOT, env=env, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_shard_of_partitions_a_scan():
    sys.path.insert(0, ROOT)
    import bench
    scan = np.arange(120_000 * 4, dtype=float).reshape(-1, 4)
    for world in (1, 2, 4, 8):
        parts = [bench.shard_of(scan, r, world) for r in range(world)]
        allp = np.concatenate(parts)
        assert len(allp) == len(scan) and np.array_equal(np.sort(allp[:, 0]), scan[:, 0])
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 32
    odd = scan[:1001]
    assert sum(len(bench.shard_of(odd, r, 3)) for r in range(3)) == 1001
    # with the scan shape: azimuth sectors (several thin ones per rank), still a partition, balanced to a few columns of 64 beams
    for world in (2, 4, 8):
        parts = [bench.shard_of(scan, r, world, n_az=1875) for r in range(world)]
        allp = np.concatenate(parts)
        assert len(allp) == len(scan) and np.array_equal(np.sort(allp[:, 0]), scan[:, 0])
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 64 * bench.SECTORS_PER_RANK


def test_reference_arm_uses_every_host_core_under_torchrun():
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm must still run on all the cores the process may use (round 1's
    N > 1 reference numbers were single-threaded) and say how many."""
    env = dict(os.environ, RANK="0", WORLD_SIZE="2", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                        "--map-points", "60000", "--cpu-fraction", "0.02"], capture_output=True, text=True, cwd=ROOT, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][0])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert f"{d['cpu_baseline']['cores']} host threads" in d["cpu_baseline"]["sample"]
