"""Shared helpers for the parity tests (test infrastructure; imports the oracle as the checker)."""
import glob
import os

import numpy as np

import tcgnn_oracle as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_sgt_files():
    return sorted(glob.glob(os.path.join(GOLDEN, "sgt_*.npz")))


def load_golden(path):
    g = np.load(path)
    return {k: g[k] for k in g.files}


def small_graphs():
    """(name, row_ptr, col_idx, n) cases used by the GPU parity tests: every golden SGT graph
    (edge cases of SURVEY.md 8a: N%16 in {0,1,15}, E=0, empty windows, hub row, unsorted rows with
    duplicated columns) plus a few larger seeded ones."""
    out = []
    for f in golden_sgt_files():
        g = load_golden(f)
        out.append((os.path.basename(f)[4:-4], g["row_pointers"], g["column_index"], int(g["num_nodes"])))
    out.append(("uniform_n5000", *orc.random_graph(5000, 200000, seed=21), 5000))
    out.append(("rmat_n20000", *orc.rmat_graph(20000, 400000, seed=22), 20000))
    return out


def sgt_arrays(rp, ci, n):
    bp, e2c, e2r, _ = orc.sgt(rp, ci, n)
    return bp, e2c, e2r


def to_dev(*arrays):
    import torch
    return [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in arrays]


def features(n, d, seed=0, kind="randn"):
    rng = np.random.default_rng(seed)
    if kind == "randn":
        return rng.standard_normal((n, d)).astype(np.float32)
    if kind == "ints":
        return rng.integers(-8, 9, size=(n, d)).astype(np.float32)
    raise ValueError(kind)


def assert_normwise(got, want, scale, tol, what=""):
    """|got - want| <= tol * (sum of |terms|) elementwise (+ a denormal-sized floor)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    bound = tol * np.asarray(scale, dtype=np.float64) + 1e-30
    bad = np.abs(got - want) > bound
    if bad.any():
        idx = np.argwhere(bad)[0]
        raise AssertionError(
            f"{what}: {int(bad.sum())} of {bad.size} elements outside {tol:g}*sum|terms|; first at {tuple(idx)}: "
            f"got {got[tuple(idx)]!r} want {want[tuple(idx)]!r} scale {np.asarray(scale)[tuple(idx)]!r}")


def assert_equal_up_to_split_windows(a, b, what="", max_rows=2 * 148 * 16, rtol=1e-5):
    """Two SpMM runs are bit-identical except in the rows of windows that are cut by a CTA slice boundary: their
    partial sums are combined with fp32 atomics / bulk reduce-adds, whose order is timing dependent when a hub
    window spans three or more CTAs.  Those rows (at most two windows per CTA) must still agree to rounding."""
    import torch
    diff = (a != b).any(dim=1)
    n_diff = int(diff.sum())
    assert n_diff <= max_rows, f"{what}: {n_diff} rows differ between two runs (more than the split windows can explain)"
    if n_diff:
        scale = torch.maximum(a[diff].abs().amax(dim=1, keepdim=True), b[diff].abs().amax(dim=1, keepdim=True))
        assert bool(((a[diff] - b[diff]).abs() <= rtol * scale + 1e-30).all()), f"{what}: split-window rows differ beyond rounding"
