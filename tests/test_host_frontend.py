"""CPU tests of the host-side mirror of the reference's Python layer: config constants, synthetic
graph generators, TCGNN_dataset attributes, gnn_conv surface (import only -- compute needs a GPU)."""
import numpy as np
import pytest
import torch


def test_config_constants():
    import config
    assert (config.BLK_H, config.BLK_W, config.WARP_SIZE) == (16, 8, 32)   # reference config.py:1-3
    assert config.func(0) == 1 and config.func(5) == 5


@pytest.mark.parametrize("kind", ["uniform", "rmat"])
def test_synthetic_graph_is_reference_format(kind):
    import graphgen
    n, target = 4000, 90000
    rp, ci = graphgen.synthetic_graph(n, target, kind=kind, seed=3)
    assert rp.dtype == torch.int32 and ci.dtype == torch.int32
    assert rp.numel() == n + 1 and int(rp[0]) == 0 and int(rp[-1]) == ci.numel()
    assert 0.99 * target <= ci.numel() <= target
    rpn, cin = rp.numpy(), ci.numpy()
    assert (np.diff(rpn) >= 0).all() and cin.min() >= 0 and cin.max() < n
    for r in (0, 1, n // 2, n - 1):                       # sorted, unique columns per row
        row = cin[rpn[r]:rpn[r + 1]]
        assert (np.diff(row) > 0).all()
    # symmetric adjacency (the reference's backward assumes A == A^T)
    from scipy.sparse import csr_matrix
    a = csr_matrix((np.ones(len(cin), np.int8), cin, rpn), shape=(n, n))
    assert (a != a.T).nnz == 0
    rp2, ci2 = graphgen.synthetic_graph(n, target, kind=kind, seed=3)
    assert torch.equal(rp, rp2) and torch.equal(ci, ci2)  # seeded


def test_dataset_synthetic_and_npz(tmp_path):
    from dataset import TCGNN_dataset
    ds = TCGNN_dataset("uniform:500:6000:1", 16, 7, load_from_txt=False, seed=0)
    assert ds.num_nodes == 500 and ds.num_features == 16 and ds.num_classes == 7
    assert ds.row_pointers.dtype == torch.int32 and ds.column_index.dtype == torch.int32
    assert ds.x.shape == (500, 16) and ds.y.shape == (500,) and int(ds.y.sum()) == 500
    assert ds.num_edges == ds.column_index.numel()
    # npz path: same CSR as scipy coo -> csr (reference dataset.py:94-104), duplicates merged
    src = np.array([0, 0, 1, 2, 2, 2]); dst = np.array([1, 1, 2, 0, 1, 1])
    f = tmp_path / "g.npz"
    np.savez(f, src_li=src, dst_li=dst, num_nodes=3)
    ds2 = TCGNN_dataset(str(f), 4, 2, load_from_txt=False)
    assert ds2.row_pointers.tolist() == [0, 1, 2, 4]
    assert ds2.column_index.tolist() == [1, 2, 0, 1]
    # name of a reference dataset that is not on disk -> the like-named synthetic workload
    ds3 = TCGNN_dataset("tcgnn-ae-graphs/cora.npz", 16, 7, load_from_txt=False, seed=0)
    assert ds3.num_nodes == 2708


def test_gnn_conv_surface():
    import gnn_conv
    for nm in ("TCGNNFunction", "TCGNNFunction_SAG", "TCGNNFunction_GIN", "TCGNNFunction_AGNN", "SAG", "GCNConv",
               "GINConv", "AGNNConv", "gen_test_tensor"):
        assert hasattr(gnn_conv, nm)
    assert gnn_conv.n_heads == 1
    conv = gnn_conv.AGNNConv(8, 4)
    assert conv.weights.shape == (8, 4) and conv.attention_w.shape == (1, 1)
    t = gnn_conv.gen_test_tensor(torch.zeros(5, 3))
    assert t.tolist() == [[float(i)] * 3 for i in range(5)]


REFERENCE = "/root/reference"


@pytest.mark.skipif(not __import__("os").path.isdir(REFERENCE), reason="reference checkout not present (GPU box)")
def test_reference_python_layer_binds_to_our_module():
    """Drop-in boundary, checked against the reference's OWN Python sources (build container only): every
    `TCGNN.<name>` that gnn_conv.py / main_tcgnn.py use exists in our extension module, and the reference's
    gnn_conv.py imports and builds its layers on top of it unchanged."""
    import importlib.util
    import os
    import re
    import sys
    import TCGNN
    used = set()
    for f in ("gnn_conv.py", "main_tcgnn.py"):
        used |= set(re.findall(r"\bTCGNN\.(\w+)\s*\(", open(os.path.join(REFERENCE, f)).read()))
    assert used >= {"forward", "forward_ef", "forward_AGNN", "preprocess"}
    missing = [n for n in sorted(used) if not hasattr(TCGNN, n)]
    assert not missing, f"reference calls TCGNN.{missing} which our module does not export"
    for n in ("preprocess_gpu", "backward", "backward_ef", "SDDMM_forward"):      # TCGNN.cpp:260-272 + north star alias
        assert hasattr(TCGNN, n)
    spec = importlib.util.spec_from_file_location("reference_gnn_conv", os.path.join(REFERENCE, "gnn_conv.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules.setdefault("TCGNN", TCGNN)
    spec.loader.exec_module(mod)                       # `import TCGNN` inside resolves to ours
    assert mod.TCGNN is TCGNN
    for cls in ("GCNConv", "GINConv", "AGNNConv", "SAG"):
        assert hasattr(mod, cls)
    conv = mod.GCNConv(8, 4)
    assert tuple(conv.weights.shape) == (8, 4)
    # same positional call contract as ours: CPU tensors must be rejected by the op with a RuntimeError
    # (CHECK_INPUT in the reference, TCGNN.cpp:54-56), not crash
    x = torch.zeros(4, 8)
    i = torch.zeros(5, dtype=torch.int32)
    with pytest.raises(RuntimeError):
        TCGNN.forward(x, i, i, i, i, i)


def test_reordering_is_an_isomorphism_and_never_needs_more_blocks_on_skewed_graphs():
    """reorder.py: the relabelled CSR is P A P^T (same edge multiset under the permutation, rows sorted), SpMM
    commutes with it (oracle), and on an R-MAT graph every order needs fewer TC blocks than the natural one; the
    block counter agrees with the SGT oracle's blockPartition sum."""
    import reorder
    import tcgnn_oracle as orc
    n = 6000
    rp, ci = orc.rmat_graph(n, 200000, seed=3)
    t_rp, t_ci = torch.from_numpy(rp), torch.from_numpy(ci)
    bp, _, _, _ = orc.sgt(rp, ci, n)
    base = reorder.count_tc_blocks(t_rp, t_ci)
    assert base == int(bp.sum())
    assert reorder.naive_tc_blocks(t_rp, t_ci) >= base
    x = np.random.default_rng(0).integers(-3, 4, size=(n, 8)).astype(np.float32)
    y = orc.spmm(x, rp, ci)
    for method in ("degree", "minhash", "hub"):
        rp2, ci2, perm, rep = reorder.reorder_graph(t_rp, t_ci, method)
        p = perm.numpy()
        assert sorted(p.tolist()) == list(range(n))
        assert rep["tc_blocks_before"] == base and rep["tc_blocks_after"] < base
        assert rep["tc_blocks_after"] == int(orc.sgt(rp2.numpy(), ci2.numpy(), n)[0].sum())
        # rows of the relabelled graph are sorted and duplicate-free, like scipy's CSR
        r2, c2 = rp2.numpy(), ci2.numpy()
        rows2 = np.repeat(np.arange(n), np.diff(r2))
        assert np.all((np.diff(c2) > 0) | (np.diff(rows2) > 0))
        # (P A P^T)(P x) = P (A x)
        assert np.array_equal(orc.spmm(x[p], r2, c2), y[p])


def test_dataset_reorder_option(tmp_path):
    from dataset import TCGNN_dataset
    ds0 = TCGNN_dataset("rmat:3000:60000:1", 8, 3, load_from_txt=False, seed=1)
    ds1 = TCGNN_dataset("rmat:3000:60000:1", 8, 3, load_from_txt=False, seed=1, reorder="hub")
    assert ds0.perm is None and ds0.reorder_report is None and not ds0.reorder_flag
    assert ds1.reorder_flag and ds1.perm.numel() == 3000
    assert ds1.column_index.numel() == ds0.column_index.numel()
    assert ds1.reorder_report["tc_blocks_after"] <= ds1.reorder_report["tc_blocks_before"]
