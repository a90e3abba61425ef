"""Host-side pieces of bench.py that need no GPU: the byte models behind the roofline keys, the measured-ceiling
interpolation and the guard that keeps the one JSON line safe once the headline is measured."""
import json
import os
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def test_algorithmic_bytes_is_the_survey_gather_model():
    # E (4 D + 4) + N (4 D + 4) for the SpMM
    assert bench.algorithmic_bytes("spmm", 10, 100, 128) == 100 * (4 * 128 + 4) + 10 * (4 * 128 + 4)


def test_compulsory_bytes_counts_every_array_once():
    n, e, d, t = 1000, 50000, 64, 7000
    spmm = bench.compulsory_bytes("spmm", n, n, e, d, t)
    assert spmm == 4 * d * 2 * n + 64 * t
    assert bench.compulsory_bytes("sddmm", n, n, e, d, t) == 4 * d * n + 64 * t + 8 * e
    assert bench.compulsory_bytes("agnn", n, n, e, d, t) == spmm + 64 * t + 8 * e


def test_ceiling_interpolates_between_sweep_points():
    curve = [{"working_set_mb": 64, "bytes_per_clk": 10000.0}, {"working_set_mb": 256, "bytes_per_clk": 6000.0}]
    assert bench.ceiling_at(curve, 10) == 10000.0
    assert bench.ceiling_at(curve, 1000) == 6000.0
    assert bench.ceiling_at(curve, 128) == pytest.approx(8000.0)       # half way in log space
    assert bench.ceiling_at(None, 128) is None


def test_workload_string_is_shared_by_both_arms():
    s = bench.workload_string("reddit-like-rmat", 5, 7, 128, "spmm", 0, "rmat")
    assert s.startswith("reddit-like-rmat: N=5 nnz=7 D=128 op=spmm")


GUARD_SCRIPT = textwrap.dedent("""
    import json, os, sys, time
    sys.path.insert(0, {root!r})
    import bench
    rank, mode = int(sys.argv[1]), sys.argv[2]
    result = {{"metric": "m", "value": 1.0, "e2e": None}}
    guard = bench.LineGuard(rank, lambda o: os.write(1, (json.dumps(o) + "\\n").encode()), result, 0.5)
    if mode == "raise":
        try:
            raise RuntimeError("variant failed")
        except Exception as exc:
            guard.bail(repr(exc))
        os.write(1, b"not reached\\n")
    elif mode == "hang":
        time.sleep(30)
        os.write(1, b"not reached\\n")
    else:
        result["e2e"] = {{"value": 2.0}}
        guard.finish()
        time.sleep(0.8)                                 # the cancelled timer must not fire
        os.write(1, (json.dumps(result) + "\\n").encode())
""")


@pytest.mark.parametrize("mode", ["raise", "hang", "ok"])
@pytest.mark.parametrize("rank", [0, 1])
def test_line_guard(tmp_path, rank, mode):
    script = tmp_path / "guard.py"
    script.write_text(GUARD_SCRIPT.format(root=ROOT))
    p = subprocess.run([sys.executable, str(script), str(rank), mode], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    if mode == "ok":
        assert len(lines) == 1 and json.loads(lines[0])["e2e"] == {"value": 2.0}
    elif rank == 0:
        assert len(lines) == 1, p.stdout               # exactly one line, the headline as it stood
        d = json.loads(lines[0])
        assert d["value"] == 1.0 and "incomplete" in d
    else:
        assert lines == []                              # the other ranks leave quietly
