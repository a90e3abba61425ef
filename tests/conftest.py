import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "tc-gnn_atc23_b200")
for p in (PKG, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "_ref"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def oracle_lib():
    """The C restatement (oracle/tcgnn_oracle.c), compiled on demand -- checker only."""
    import ctypes
    import subprocess
    src = os.path.join(ROOT, "oracle", "tcgnn_oracle.c")
    out_dir = os.path.join(ROOT, "oracle", "_build")
    out = os.path.join(out_dir, "libtcgnn_oracle.so")
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(out_dir, exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-fopenmp", "-fPIC", "-shared", src, "-o", out])
    lib = ctypes.CDLL(out)
    lib.oracle_sgt.restype = ctypes.c_int64
    return lib
