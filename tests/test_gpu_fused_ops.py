"""GPU parity of the entry points added around the three core kernels: the fused AGNN pipeline (tcgnn_agnn_f32),
tile-ordered weights, TCGNN_ACCUMULATE partial products over source panels, the pipelined host-buffer entries of all
three ops, the transposed-graph plan (directed graphs) and the exchange helpers -- each against the NumPy oracle."""
import ctypes as C

import numpy as np
import pytest

import tcgnn_oracle as orc
from _util import assert_normwise, features, sgt_arrays, small_graphs, to_dev

pytestmark = pytest.mark.gpu


def _graph(rp, ci, n):
    bp, e2c, e2r = sgt_arrays(rp, ci, n)
    return tuple(to_dev(rp, ci, bp, e2c, e2r))


@pytest.mark.parametrize("case", small_graphs(), ids=lambda c: c[0])
@pytest.mark.parametrize("d", [16, 64, 100, 256])
def test_fused_agnn_matches_oracle_layer_and_three_calls(case, d):
    """Y, the CSR-order scores and the tile-ordered attention of tcgnn_agnn_f32 against the oracle's restatement of
    gnn_conv.py:125-132 and against our own three-call sequence (bit for bit: same kernels, same order)."""
    import torch
    import TCGNN
    name, rp, ci, n = case
    if len(ci) == 0:
        pytest.skip("no edges")
    g = _graph(rp, ci, n)
    x = features(n, d, seed=7) * 0.25
    aw = np.float32(0.6)
    d_x = torch.from_numpy(x).cuda()
    d_aw = torch.full((1, 1), float(aw), device="cuda")
    y, att_tile, ef = TCGNN.forward_AGNN_fused(d_x, g[0], g[1], d_aw, g[2], g[3], g[4], True)
    ef_o = orc.sddmm(x, rp, ci)
    dup = len(np.unique(np.repeat(np.arange(n), np.diff(rp)).astype(np.int64) * n + ci)) < len(ci)
    ef_h = ef.cpu().numpy()
    if dup:   # one edge of every duplicated (row, col) pair receives the score, the others stay 0 (as in the reference)
        keep = ef_h != 0
        assert_normwise(ef_h[keep], ef_o[keep], orc.sddmm_abs(x, rp, ci)[keep], 1e-5, f"{name} fused scores")
    else:
        assert_normwise(ef_h, ef_o, orc.sddmm_abs(x, rp, ci), 1e-5, f"{name} fused scores")
    ef3 = TCGNN.forward_ef(d_x, *g)[0]
    att3 = torch.mm(ef3.unsqueeze(-1), d_aw).transpose(0, 1).contiguous()
    y3 = TCGNN.forward_AGNN(d_x, g[0], g[1], att3, g[2], g[3], g[4])[0]
    w = (ef3.cpu().numpy() * aw).astype(np.float32)
    scale = orc.spmm_abs(x, rp, ci, w)
    if not dup:
        assert torch.equal(ef, ef3)
        # identical kernels on identical operands; only the reduce-add order of windows split across CTAs may differ
        assert_normwise(y.cpu().numpy(), y3.cpu().numpy(), scale, 1e-6, f"{name} fused vs 3 calls")
        assert_normwise(y.cpu().numpy(), orc.spmm(x, rp, ci, w), scale, 1e-5, f"{name} fused Y")
    yb = TCGNN.forward_AGNN_tile(d_x, g[0], g[1], att_tile, g[2], g[3], g[4])[0]
    assert_normwise(yb.cpu().numpy(), y.cpu().numpy(), scale, 1e-6, f"{name} tile weights")


def test_fused_agnn_through_the_c_abi_exact_on_integers():
    import torch
    import tcgnn_capi as capi
    n, d = 5000, 128
    rp, ci = orc.rmat_graph(n, 150000, seed=3)
    g = _graph(rp, ci, n)
    plan = capi.Plan(*g)
    xi = np.random.default_rng(1).integers(-2, 3, size=(n, d)).astype(np.float32)
    d_x = torch.from_numpy(xi).cuda()
    aw = torch.full((1,), 2.0, device="cuda")
    y = torch.empty(n, d, device="cuda")
    ef = torch.empty(len(ci), device="cuda")
    att = torch.empty(plan.info()["pairs"], device="cuda")
    plan.agnn(d_x, aw, y, att_tile_out=att, edge_out=ef)
    ef_o = orc.sddmm(xi, rp, ci)
    assert np.array_equal(ef.cpu().numpy(), ef_o)
    assert np.array_equal(y.cpu().numpy(), orc.spmm(xi, rp, ci, (ef_o * 2).astype(np.float32)))
    # attention_w == NULL means 1.0; no outputs requested besides Y
    y1 = torch.empty(n, d, device="cuda")
    plan.agnn(d_x, None, y1)
    assert np.array_equal(y1.cpu().numpy(), orc.spmm(xi, rp, ci, ef_o))
    # tile-ordered weights straight into tcgnn_spmm_f32_ex
    y2 = torch.empty(n, d, device="cuda")
    plan.spmm_ex(d_x, y2, edge_weight=att, flags=capi.W_TILE_ORDER)
    assert torch.equal(y2, y)
    plan.close()


def test_accumulate_mode_adds_partial_products_over_column_ranges():
    """Y = sum_p A[:, cols of p] X[p] with TCGNN_ACCUMULATE == the one-shot product (integer data: exact)."""
    import torch
    import TCGNN
    n, d = 7001, 96
    rp, ci = orc.rmat_graph(n, 250000, seed=9)
    xi = features(n, d, seed=2, kind="ints")
    want = orc.spmm(xi, rp, ci)
    rows = np.repeat(np.arange(n), np.diff(rp))
    y = None
    bounds = [0, 1500, 1504, 4000, n]
    for b0, b1 in zip(bounds, bounds[1:]):
        m = (ci >= b0) & (ci < b1)
        sub_rp = np.zeros(n + 1, dtype=np.int32)
        np.cumsum(np.bincount(rows[m], minlength=n), out=sub_rp[1:])
        sub_ci = (ci[m] - b0).astype(np.int32)
        bp, e2c, e2r, _ = orc.sgt(sub_rp, sub_ci, n)
        g = to_dev(sub_rp, sub_ci, bp, e2c, e2r)
        xs = torch.from_numpy(xi[b0:b1]).cuda()
        if y is None:
            y = TCGNN.source_forward(xs, *g)[0]
        else:
            out = TCGNN.source_forward(xs, *g, accumulate_into=y)[0]
            assert out.data_ptr() == y.data_ptr()
    assert np.array_equal(y.cpu().numpy(), want)


@pytest.mark.parametrize("n,e,d,kind", [(70000, 1500000, 64, "rmat"), (70000, 1500000, 132, "uniform"),
                                        (3000, 40000, 32, "rmat")])
def test_host_buffer_entries_match_resident_ops(n, e, d, kind):
    """tcgnn_{spmm,sddmm,agnn}_f32_host (pinned host X in, results out; SpMM pipelined over column / row chunks when
    the graph is large enough) == the device-resident operators on the same data."""
    import torch
    import TCGNN
    rp, ci = (orc.rmat_graph if kind == "rmat" else orc.random_graph)(n, e, seed=11)
    g = _graph(rp, ci, n)
    xi = features(n, d, seed=4, kind="ints")
    x_host = torch.from_numpy(xi).pin_memory()
    d_x = x_host.cuda()
    y_host = TCGNN.forward_host(x_host, *g)
    assert np.array_equal(y_host.numpy(), orc.spmm(xi, rp, ci))
    # a second call re-uses the staging buffers and the sub-plans; caller-provided result buffer, async
    y2 = torch.empty(n, d).pin_memory()
    TCGNN.forward_host(x_host, *g, y_host=y2, sync=False)
    torch.cuda.synchronize()
    assert torch.equal(y2, y_host)
    xf = torch.from_numpy(features(n, d, seed=5)).pin_memory()
    yr = TCGNN.forward(xf.cuda(), *g)[0].cpu()
    yh = TCGNN.forward_host(xf, *g)
    assert_normwise(yh.numpy(), yr.numpy(), orc.spmm_abs(xf.numpy(), rp, ci), 1e-5, "host SpMM on normal data")
    e_host = TCGNN.forward_ef_host(x_host, *g)
    assert torch.equal(e_host, TCGNN.forward_ef(d_x, *g)[0].cpu())
    aw = torch.full((1, 1), 0.5, device="cuda")
    ya = TCGNN.forward_AGNN_host(x_host, g[0], g[1], aw, g[2], g[3], g[4])
    yd = TCGNN.forward_AGNN_fused(d_x, g[0], g[1], aw, g[2], g[3], g[4], False)[0].cpu()
    assert torch.equal(ya, yd)


def test_transposed_plan_gives_the_directed_backward():
    """dX = A^T dY on a NON-symmetric graph (the reference's backward re-uses A, gnn_conv.py:76-85): SpMM and weighted
    SpMM over the device-built transposed graph against the oracle on scipy's A^T."""
    import scipy.sparse as sp
    import torch
    import TCGNN
    n, d = 4003, 48
    rp, ci = _directed(n, 120000)
    a = sp.csr_matrix((np.ones(len(ci), dtype=np.float32), ci, rp), shape=(n, n))
    assert (a != a.T).nnz > 0
    g = _graph(rp, ci, n)
    dy = features(n, d, seed=6, kind="ints")
    d_dy = torch.from_numpy(dy).cuda()
    got = TCGNN.backward_T(d_dy, *g)[0].cpu().numpy()
    at = a.T.tocsr()
    at.sort_indices()
    assert np.array_equal(got, orc.spmm(dy, at.indptr.astype(np.int32), at.indices.astype(np.int32)))
    assert not np.array_equal(got, TCGNN.backward(d_dy, *g)[0].cpu().numpy())      # A != A^T here
    w = np.random.default_rng(3).integers(-3, 4, size=len(ci)).astype(np.float32)
    aw = sp.csr_matrix((w, ci, rp), shape=(n, n)).T.tocsr()
    aw.sort_indices()
    got_w = TCGNN.backward_T_AGNN(d_dy, g[0], g[1], torch.from_numpy(w).cuda().reshape(1, -1), g[2], g[3], g[4])[0]
    want_w = np.asarray(aw.astype(np.float64) @ dy.astype(np.float64))
    assert np.array_equal(got_w.cpu().numpy().astype(np.float64), want_w)
    # the raw transpose: same edge multiset, edge_map carries each edge back to its CSR slot in A
    rp_t, ci_t, map_t = TCGNN.csr_transpose(g[0], g[1])
    assert np.array_equal(rp_t.cpu().numpy(), at.indptr)
    rows_t = np.repeat(np.arange(n), np.diff(at.indptr))
    m = map_t.cpu().numpy()
    assert np.array_equal(np.sort(m), np.arange(len(ci)))
    rows_a = np.repeat(np.arange(n), np.diff(rp))
    assert np.array_equal(ci[m], rows_t) and np.array_equal(rows_a[m], ci_t.cpu().numpy())


def _directed(n, e):
    rng = np.random.default_rng(17)
    src = rng.integers(0, n, e)
    dst = (rng.integers(0, n, e) ** 2 // n).astype(np.int64)      # skewed targets
    key = np.unique(src.astype(np.int64) * n + dst)
    rows, cols = key // n, (key % n).astype(np.int32)
    rp = np.zeros(n + 1, dtype=np.int32)
    np.cumsum(np.bincount(rows, minlength=n), out=rp[1:])
    return rp, cols


def test_autograd_layers_on_a_directed_graph():
    """gnn_conv.set_assume_symmetric(False): GCN dX / dW on a directed graph against dense autograd."""
    import torch
    import gnn_conv
    n, din, dout = 900, 24, 16
    rp, ci = _directed(n, 20000)
    g = _graph(rp, ci, n)
    a = torch.zeros(n, n, dtype=torch.float64)
    rows = np.repeat(np.arange(n), np.diff(rp))
    a[torch.from_numpy(rows), torch.from_numpy(ci.astype(np.int64))] = 1.0
    gen = torch.Generator().manual_seed(5)
    x = torch.randint(-2, 3, (n, din), generator=gen).float()
    w = torch.randint(-2, 3, (din, dout), generator=gen).float()
    gy = torch.randint(-2, 3, (n, dout), generator=gen).float()
    gnn_conv.set_assume_symmetric(False)
    try:
        xc = x.cuda().requires_grad_(True)
        wc = w.cuda().requires_grad_(True)
        y = gnn_conv.TCGNNFunction.apply(xc, wc, *g)
        y.backward(gy.cuda())
    finally:
        gnn_conv.set_assume_symmetric(True)
    xr = x.double().requires_grad_(True)
    wr = w.double().requires_grad_(True)
    (a @ (xr @ wr)).backward(gy.double())
    assert torch.equal(y.detach().cpu().double(), (a @ (x.double() @ w.double())))
    assert torch.equal(xc.grad.cpu().double(), xr.grad)
    assert torch.equal(wc.grad.cpu().double(), wr.grad)


def test_weighted_spmm_with_duplicated_pairs():
    """Duplicated (row, col) entries: the reference's kernel lets one of the duplicates' attention values win
    (TCGNN_kernel.cu:529, a benign race); ours picks one deterministically.  With one-hot features Y[i, j] IS the
    weight used for pair (i, j): it must be tf32(weight of one of its duplicates)."""
    import torch
    import TCGNN
    from _util import golden_sgt_files, load_golden
    f = [p for p in golden_sgt_files() if "unsorted_dups" in p][0]
    gd = load_golden(f)
    rp, ci, n = gd["row_pointers"], gd["column_index"], int(gd["num_nodes"])
    g = to_dev(rp, ci, gd["blockPartition"], gd["edgeToColumn"], gd["edgeToRow"])
    rows = np.repeat(np.arange(n), np.diff(rp))
    w = np.random.default_rng(8).standard_normal(len(ci)).astype(np.float32)
    x = np.eye(n, dtype=np.float32)
    y = TCGNN.forward_AGNN(torch.from_numpy(x).cuda(), g[0], g[1], torch.from_numpy(w).cuda().reshape(1, -1), *g[2:])[0]
    y = y.cpu().numpy()
    wr = orc.tf32_rna(w)
    allowed = {}
    for e in range(len(ci)):
        allowed.setdefault((int(rows[e]), int(ci[e])), set()).add(float(wr[e]))
    assert any(len(v) > 1 for v in allowed.values()), "fixture must contain duplicated pairs"
    nz = np.argwhere(y != 0)
    for i, j in nz:
        assert float(y[i, j]) in allowed[(int(i), int(j))]
    assert len(nz) <= len(allowed)
    # same weight on every duplicate: the sum over DISTINCT pairs, exactly (pattern semantics of the SGT)
    pair_w = {k: np.float32(hash(k) % 7 - 3) for k in allowed}
    w2 = np.array([pair_w[(int(rows[e]), int(ci[e]))] for e in range(len(ci))], dtype=np.float32)
    xi = features(n, 32, seed=1, kind="ints")
    y2 = TCGNN.forward_AGNN(torch.from_numpy(xi).cuda(), g[0], g[1], torch.from_numpy(w2).cuda().reshape(1, -1), *g[2:])[0]
    want = np.zeros((n, 32), dtype=np.float64)
    for (i, j), v in pair_w.items():
        want[i] += float(v) * xi[j]
    assert np.array_equal(y2.cpu().numpy().astype(np.float64), want)


def test_exchange_helpers():
    import torch
    import TCGNN
    src = torch.randn(1000, 64, device="cuda")
    rows = torch.randint(0, 1000, (333,), device="cuda", dtype=torch.int32)
    dst = torch.zeros(400, 64, device="cuda")
    TCGNN.gather_rows(src, rows, dst)
    assert torch.equal(dst[:333], src[rows.long()]) and float(dst[333:].abs().max()) == 0.0
    flags = torch.tensor([5, 0], dtype=torch.int32, device="cuda")
    err = torch.zeros(1, dtype=torch.int32, device="cuda")
    TCGNN.stream_wait_flag(flags, 0, 5, 1000, err)         # already satisfied
    TCGNN.stream_wait_flag(flags, 0, 3, 1000, err)         # a later value also satisfies
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    TCGNN.stream_wait_flag(flags, 1, 1, 50, err)           # never set: times out instead of hanging, reports it
    torch.cuda.synchronize()
    assert int(err.item()) == 1
    # set from another stream while the first one waits
    err.zero_()
    s2 = torch.cuda.Stream()
    TCGNN.stream_wait_flag(flags, 1, 7, 5000, err)
    with torch.cuda.stream(s2):
        flags[1:2].copy_(torch.tensor([7], dtype=torch.int32, device="cuda"))
    torch.cuda.synchronize()
    assert int(err.item()) == 0


def test_sddmm_rejects_detached_plans_and_prerounded_ragged_widths():
    """TCGNN_X_IS_TF32 with dim % 4 != 0 must not contract over the caller's extra columns: X is a
    column slice of a wider matrix whose remaining columns are non-zero."""
    import torch
    import tcgnn_capi as capi
    n, d = 3000, 30
    rp, ci = orc.random_graph(n, 50000, seed=23)
    g = _graph(rp, ci, n)
    wide = features(n, 32, seed=3)
    wide_r = orc.tf32_rna(wide)
    d_wide = torch.from_numpy(wide_r).cuda()                 # rounded already, 16-byte rows, columns 30..31 non-zero
    plan = capi.Plan(*g)
    out = torch.empty(len(ci), device="cuda")
    capi.check(capi.lib().tcgnn_sddmm_f32_ex(plan._h, capi._ptr(d_wide), 32, capi._ptr(out), d, capi.X_IS_TF32,
                                             capi._stream()), "sddmm_ex")
    x = np.ascontiguousarray(wide[:, :d])
    assert_normwise(out.cpu().numpy(), orc.sddmm(x, rp, ci), orc.sddmm_abs(x, rp, ci), 1e-5, "ragged pre-rounded SDDMM")
    plan.close()
