"""tools/verify_thirdparty.cpp — the closing procedure for the "unpinned against the real Eigen / Sophus / tsl::robin_map" caveat:
an integrator builds it against the real libraries and replays tests/golden/thirdparty_vectors.txt.  Here (no such libraries) it
is built against the stand-in headers of oracle/shim, where it must pass, and the committed vectors must be what the oracle gives."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VECTORS = os.path.join(ROOT, "tests", "golden", "thirdparty_vectors.txt")


def test_committed_vectors_are_the_oracles(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_thirdparty_vectors as gen
    committed = open(VECTORS).read()
    gen.OUT = str(tmp_path / "v.txt")
    gen.main()
    assert open(gen.OUT).read() == committed  # deterministic generator, committed result


def test_selfcheck_builds_and_passes_against_the_stand_ins(tmp_path):
    exe = str(tmp_path / "verify_thirdparty")
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "oracle", "shim"), "-I", os.path.join(ROOT, "oracle"),
                           os.path.join(ROOT, "tools", "verify_thirdparty.cpp"), "-o", exe])
    r = subprocess.run([exe, VECTORS], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "checks passed" in r.stdout
    # a corrupted vector must be caught (the checker really compares)
    bad = tmp_path / "bad.txt"
    lines = open(VECTORS).read().splitlines()
    i = next(k for k, l in enumerate(lines) if l.startswith("ROBIN growth"))
    n_ops = int(lines[i].split()[2])
    order = lines[i + 1 + n_ops].split()
    order[0], order[1] = order[1], order[0]
    lines[i + 1 + n_ops] = " ".join(order)
    bad.write_text("\n".join(lines) + "\n")
    r = subprocess.run([exe, str(bad)], capture_output=True, text=True)
    assert r.returncode == 1 and "robin_map iteration order, case growth" in r.stderr
