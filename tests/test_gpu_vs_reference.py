"""Parity against the REFERENCE'S OWN CUDA KERNELS run on the same GPU (oracle/_ref/TCGNN_ref*.so,
built unmodified from /root/reference/TCGNN_conv by oracle/build_ref.sh; it carries sm_100 SASS
and travels with the snapshot).  This pins the SpMM / SDDMM oracle restatement and the new kernels
to the reference at once.  Restricted to the region where the reference is defined
(SURVEY.md 8a): D % 16 == 0 and D <= 128, N % 16 == 0 (its last window stores 16 full rows),
E <= 2^24 for SDDMM.  Tolerance 1e-3 * sum|terms| (north star); in practice the two agree to
accumulation order (~1e-6)."""
import os

import numpy as np
import pytest

import tcgnn_oracle as orc
from _util import assert_normwise, features, sgt_arrays, to_dev

pytestmark = pytest.mark.gpu

try:
    import TCGNN_ref  # noqa: F401
    HAVE_REF = True
except Exception:  # pragma: no cover
    HAVE_REF = False

needs_ref = pytest.mark.skipif(not HAVE_REF, reason="oracle/_ref/TCGNN_ref*.so not built (oracle/build_ref.sh)")

CASES = [("uniform", 4096, 120000, 0), ("rmat", 8192, 200000, 1), ("citeseer_like", 3328, 9464, 2)]


def graph(kind, n, e, seed):
    return orc.rmat_graph(n, e, seed=seed) if kind == "rmat" else orc.random_graph(n, e, seed=seed)


@needs_ref
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("dim", [16, 64, 128])
def test_forward_matches_reference_kernel(case, dim):
    import torch
    import TCGNN
    import TCGNN_ref
    kind, n, e, seed = case
    rp, ci = graph(kind, n, e, seed)
    bp, e2c, e2r = sgt_arrays(rp, ci, n)
    x = features(n, dim, seed=seed + 40)
    d_rp, d_ci, d_bp, d_e2c, d_e2r, d_x = to_dev(rp, ci, bp, e2c, e2r, x)
    ref = TCGNN_ref.forward(d_x, d_rp, d_ci, d_bp, d_e2c, d_e2r)[0]
    torch.cuda.synchronize()
    new = TCGNN.forward(d_x, d_rp, d_ci, d_bp, d_e2c, d_e2r)[0]
    torch.cuda.synchronize()
    scale = orc.spmm_abs(x, rp, ci)
    assert_normwise(new.cpu().numpy(), ref.cpu().numpy(), scale, 1e-3, "new kernel vs reference kernel")
    assert_normwise(orc.spmm(x, rp, ci), ref.cpu().numpy(), scale, 1e-5, "oracle vs reference kernel")


@needs_ref
@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_forward_agnn_and_ef_match_reference_kernels(case):
    import torch
    import TCGNN
    import TCGNN_ref
    kind, n, e, seed = case
    rp, ci = graph(kind, n, e, seed)
    bp, e2c, e2r = sgt_arrays(rp, ci, n)
    x = features(n, 64, seed=seed + 50)
    d_rp, d_ci, d_bp, d_e2c, d_e2r, d_x = to_dev(rp, ci, bp, e2c, e2r, x)
    ef_ref = TCGNN_ref.forward_ef(d_x, d_rp, d_ci, d_bp, d_e2c, d_e2r)[0]
    torch.cuda.synchronize()
    ef_new = TCGNN.forward_ef(d_x, d_rp, d_ci, d_bp, d_e2c, d_e2r)[0]
    torch.cuda.synchronize()
    sc = orc.sddmm_abs(x, rp, ci)
    assert_normwise(ef_new.cpu().numpy(), ef_ref.cpu().numpy(), sc, 1e-3, "SDDMM new vs reference kernel")
    assert_normwise(orc.sddmm(x, rp, ci), ef_ref.cpu().numpy(), sc, 1e-5, "SDDMM oracle vs reference kernel")
    att = (ef_ref * 0.37).reshape(1, -1).contiguous()
    y_ref = TCGNN_ref.forward_AGNN(d_x, d_rp, d_ci, att, d_bp, d_e2c, d_e2r)[0]
    torch.cuda.synchronize()
    y_new = TCGNN.forward_AGNN(d_x, d_rp, d_ci, att, d_bp, d_e2c, d_e2r)[0]
    torch.cuda.synchronize()
    w = att[0].cpu().numpy()
    sc = orc.spmm_abs(x, rp, ci, w)
    assert_normwise(y_new.cpu().numpy(), y_ref.cpu().numpy(), sc, 1e-3, "weighted SpMM new vs reference kernel")
    assert_normwise(orc.spmm(x, rp, ci, w), y_ref.cpu().numpy(), sc, 1e-5, "weighted SpMM oracle vs reference kernel")


@needs_ref
def test_preprocess_matches_reference_preprocess_live():
    import torch
    import TCGNN
    import TCGNN_ref
    n = 6000
    rp, ci = orc.rmat_graph(n, 90000, seed=9)
    outs = []
    for mod in (TCGNN_ref, TCGNN):
        bp = torch.zeros((n + 15) // 16 + 1, dtype=torch.int32)
        e2c = torch.zeros(len(ci), dtype=torch.int32)
        e2r = torch.zeros(len(ci), dtype=torch.int32)
        mod.preprocess(torch.from_numpy(ci.copy()), torch.from_numpy(rp.copy()), n, 16, 8, bp, e2c, e2r)
        outs.append((bp[:-1].numpy().copy(), e2c.numpy().copy(), e2r.numpy().copy()))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
