"""The reference's OWN driver, unmodified, on top of the new `TCGNN` module on the GPU.

oracle/build_ref.sh stages /root/reference/{main_tcgnn.py, gnn_conv.py, dataset.py, config.py} byte for byte into
oracle/_ref/driver/ (git-ignored like the compiled reference module; /root/reference itself does not exist on the GPU
box).  The test writes a synthetic `tcgnn-ae-graphs/<name>.npz` in the reference's format (dataset.py:69-80: src_li,
dst_li, num_nodes), runs `python main_tcgnn.py --dataset <name> ...` exactly as 0_run_tcgnn_model.sh /
2_tcgnn_single_kernel.py do, with only PYTHONPATH pointing at the new module, and checks the log lines the
reference's scrapers parse (1_log2csv.py).  The edge list contains duplicated pairs on purpose: scipy merges them, so
main_tcgnn.py:44-46 allocates edgeToColumn / edgeToRow longer than column_index."""
import hashlib
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "driver")


def _run(tmp_path, *args):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.path.join(ROOT, "tc-gnn_atc23_b200") + os.pathsep + env.get("PYTHONPATH", "")
    res = subprocess.run([sys.executable, os.path.join(DRIVER, "main_tcgnn.py"), *args], cwd=tmp_path, env=env,
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-4000:]
    return res.stdout


@pytest.fixture(scope="module")
def workdir(tmp_path_factory):
    if not os.path.exists(os.path.join(DRIVER, "main_tcgnn.py")):
        pytest.skip("oracle/_ref/driver not staged (run oracle/build_ref.sh where /root/reference exists)")
    sums = {}
    with open(os.path.join(DRIVER, "SHA256SUMS")) as fh:
        for line in fh:
            h, name = line.split()
            sums[name] = h
    for name, h in sums.items():      # the staged files are the files that were hashed when they were copied
        with open(os.path.join(DRIVER, name), "rb") as fh:
            assert hashlib.sha256(fh.read()).hexdigest() == h, f"{name} was modified after staging"
    d = tmp_path_factory.mktemp("refdriver")
    os.makedirs(d / "tcgnn-ae-graphs")
    rng = np.random.default_rng(3)
    n, m = 3327, 4800
    src = rng.integers(0, n, m)
    dst = rng.integers(0, n, m)
    src, dst = np.concatenate([src, dst, src[:200]]), np.concatenate([dst, src, dst[:200]])   # symmetric + duplicates
    np.savez(d / "tcgnn-ae-graphs" / "synth.npz", src_li=src, dst_li=dst, num_nodes=n)
    return d


def test_reference_main_gcn_runs_unmodified(workdir):
    out = _run(workdir, "--dataset", "synth", "--dim", "16", "--hidden", "16", "--classes", "7", "--epochs", "5",
               "--model", "gcn")
    assert "TC_Blocks:" in out and "Exp_Edges:" in out
    assert "Prep. (ms):" in out and "Train (ms):" in out


def test_reference_main_agnn_runs_unmodified(workdir):
    out = _run(workdir, "--dataset", "synth", "--dim", "32", "--hidden", "32", "--classes", "7", "--epochs", "3",
               "--num_layers", "4", "--model", "agnn")
    assert "Train (ms):" in out


def test_reference_single_kernel_profile_runs_unmodified(workdir):
    """2_tcgnn_single_kernel.py's command line (BASELINE.json configs[1]: dim = hidden = 16, SAG.profile)."""
    out = _run(workdir, "--dataset", "synth", "--dim", "16", "--hidden", "16", "--classes", "7", "--single_kernel")
    assert "=> SAG profiling avg (ms):" in out
