"""SDDMM parity on the GPU (tcgnn_sddmm_f32 / TCGNN.forward_ef / TCGNN.SDDMM_forward).
Tolerances as in test_gpu_spmm.py: 1e-5*sum|terms| vs the TF32-emulating oracle (accumulation
order only), 1e-3 (2e-3: both operands are rounded) vs fp64 true math."""
import numpy as np
import pytest

import tcgnn_oracle as orc
from _util import assert_normwise, features, small_graphs, sgt_arrays, to_dev

pytestmark = pytest.mark.gpu

GRAPHS = small_graphs()
IDS = [g[0] for g in GRAPHS]


def run_sddmm(rp, ci, n, x, via="capi"):
    import torch
    import tcgnn_capi
    bp, e2c, e2r = sgt_arrays(rp, ci, n)
    d_rp, d_ci, d_bp, d_e2c, d_e2r, d_x = to_dev(rp, ci, bp, e2c, e2r, x)
    if via == "capi":
        plan = tcgnn_capi.Plan(d_rp, d_ci, d_bp, d_e2c, d_e2r)
        out = torch.full((len(ci),), float("nan"), device="cuda")
        plan.sddmm(d_x, out)
        torch.cuda.synchronize()
        res = out.cpu().numpy()
        plan.close()
        return res
    import TCGNN
    out = TCGNN.SDDMM_forward(d_x, d_rp, d_ci, d_bp, d_e2c, d_e2r)[0]
    torch.cuda.synchronize()
    return out.cpu().numpy()


@pytest.mark.parametrize("graph", GRAPHS, ids=IDS)
@pytest.mark.parametrize("dim", [32, 128])
def test_sddmm_matches_oracle(graph, dim):
    name, rp, ci, n = graph
    if name == "unsorted_dups_n200":
        pytest.skip("duplicated (row, col) pairs: only one edge of a pair is written (reference: last writer wins)")
    x = features(n, dim, seed=13)
    got = run_sddmm(rp, ci, n, x)
    scale = orc.sddmm_abs(x, rp, ci)
    assert_normwise(got, orc.sddmm(x, rp, ci), scale, 1e-5, f"{name} D={dim} vs tf32 oracle")
    assert_normwise(got, orc.sddmm(x, rp, ci, tf32=False, dtype=np.float64), scale, 2e-3, f"{name} D={dim} vs fp64")


@pytest.mark.parametrize("dim", [1, 8, 12, 22, 64, 100, 256, 300])
def test_sddmm_feature_widths(dim):
    rp, ci = orc.rmat_graph(3000, 60000, seed=5)
    x = features(3000, dim, seed=14)
    got = run_sddmm(rp, ci, 3000, x, via="module")
    assert_normwise(got, orc.sddmm(x, rp, ci), orc.sddmm_abs(x, rp, ci), 1e-5, f"D={dim}")


def test_sddmm_exact_on_integers():
    rp, ci = orc.rmat_graph(6000, 150000, seed=15)
    x = features(6000, 64, seed=16, kind="ints")
    got = run_sddmm(rp, ci, 6000, x)
    assert np.array_equal(got, orc.sddmm(x, rp, ci))


def test_sddmm_duplicate_pairs_write_one_edge_per_pair():
    from _util import load_golden, GOLDEN
    g = load_golden(f"{GOLDEN}/sgt_unsorted_dups_n200.npz")
    rp, ci, n = g["row_pointers"], g["column_index"], 200
    x = features(n, 32, seed=17, kind="ints")
    got = run_sddmm(rp, ci, n, x)
    want = orc.sddmm(x, rp, ci)
    rows = np.repeat(np.arange(n), np.diff(rp))
    key = rows.astype(np.int64) * n + ci
    for k in np.unique(key):
        idx = np.nonzero(key == k)[0]
        vals = got[idx]
        assert (vals == want[idx[0]]).sum() >= 1           # one edge of the pair carries the value
        assert np.all((vals == want[idx[0]]) | (vals == 0))  # the others stay zero
