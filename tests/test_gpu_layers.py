"""The autograd wrappers / conv layers (gnn_conv.py) on the GPU against the oracle's layer
restatements (reference gnn_conv.py:54-158) and against the hand-derived gradients."""
import numpy as np
import pytest

import tcgnn_oracle as orc
from _util import assert_normwise, features, sgt_arrays, to_dev

pytestmark = pytest.mark.gpu


def setup(n=2500, e=50000, seed=61):
    rp, ci = orc.random_graph(n, e, seed=seed)      # symmetric: backward re-uses the forward CSR
    bp, e2c, e2r = sgt_arrays(rp, ci, n)
    return rp, ci, tuple(to_dev(rp, ci, bp, e2c, e2r))


def test_gcn_layer_forward_backward():
    import torch
    import gnn_conv
    rp, ci, g = setup()
    n = len(rp) - 1
    x = features(n, 40, seed=62)
    w = features(40, 32, seed=63)
    X = torch.from_numpy(x).cuda().requires_grad_(True)
    conv = gnn_conv.GCNConv(40, 32).cuda()
    conv.weights.data.copy_(torch.from_numpy(w))
    y = conv(X, *g)
    xw = x @ w
    assert_normwise(y.detach().cpu().numpy(), orc.spmm(xw, rp, ci), orc.spmm_abs(xw, rp, ci) + 1e-3, 1e-4, "GCN fwd")
    dy = features(n, 32, seed=64)
    y.backward(torch.from_numpy(dy).cuda())
    dxp = orc.spmm(dy, rp, ci)                       # reference gnn_conv.py:80-84
    sc = orc.spmm_abs(dy, rp, ci)
    assert_normwise(X.grad.cpu().numpy(), dxp @ w.T, sc @ np.abs(w.T) + 1e-3, 1e-4, "GCN dX")
    assert_normwise(conv.weights.grad.cpu().numpy(), x.T @ dxp, np.abs(x.T) @ sc + 1e-3, 1e-4, "GCN dW")


def test_gin_layer_and_sag():
    import torch
    import gnn_conv
    rp, ci, g = setup(seed=65)
    n = len(rp) - 1
    x = features(n, 48, seed=66)
    w = features(48, 16, seed=67)
    X = torch.from_numpy(x).cuda().requires_grad_(True)
    conv = gnn_conv.GINConv(48, 16).cuda()
    conv.weights.data.copy_(torch.from_numpy(w))
    y = conv(X, *g)
    ax = orc.spmm(x, rp, ci)
    assert_normwise(y.detach().cpu().numpy(), ax @ w, orc.spmm_abs(x, rp, ci) @ np.abs(w) + 1e-3, 1e-4, "GIN fwd")
    y.sum().backward()
    assert X.grad.shape == X.shape and conv.weights.grad.shape == conv.weights.shape
    sag = gnn_conv.SAG(*g)
    X2 = torch.from_numpy(x).cuda().requires_grad_(True)
    out = sag(X2)
    assert_normwise(out.detach().cpu().numpy(), ax, orc.spmm_abs(x, rp, ci), 1e-5, "SAG fwd")
    out.backward(torch.ones_like(out))
    deg = np.diff(rp).astype(np.float32)
    assert np.array_equal(X2.grad.cpu().numpy(), np.repeat(deg[:, None], 48, 1))   # A^T 1 = degree (A symmetric)


def test_agnn_layer_forward_backward_shapes_and_values():
    import torch
    import gnn_conv
    rp, ci, g = setup(n=1800, e=30000, seed=68)
    n = len(rp) - 1
    x = features(n, 24, seed=69)
    conv = gnn_conv.AGNNConv(24, 16).cuda()
    w = conv.weights.detach().cpu().numpy()
    aw = conv.attention_w.detach().cpu().numpy()
    X = torch.from_numpy(x).cuda().requires_grad_(True)
    y = conv(X, *g)
    y_o, ef_o, att_o = orc.agnn_layer_forward(x, w, aw, rp, ci)
    xp = x @ w
    sc = orc.spmm_abs(xp, rp, ci, att_o[0])
    assert_normwise(y.detach().cpu().numpy(), y_o, sc + 1e-3, 2e-3, "AGNN fwd")
    y.backward(torch.ones_like(y))
    assert X.grad.shape == X.shape
    assert conv.weights.grad.shape == conv.weights.shape
    assert conv.attention_w.grad.shape == conv.attention_w.shape      # reference gnn_conv.py:150-155


def test_main_tcgnn_single_kernel_and_training_smoke(capfd):
    """main_tcgnn.py end to end on a synthetic cora-sized graph: SGT on host threads, the reference's
    log lines, a few GCN and AGNN epochs."""
    import main_tcgnn
    assert main_tcgnn.main(["--dataset", "cora", "--dim", "16", "--hidden", "16", "--single_kernel", "--seed", "0"]) == 0
    out = capfd.readouterr().out
    assert "Prep. (ms):" in out and "=> SAG profiling avg (ms):" in out and "TC_Blocks:" in out
    assert main_tcgnn.main(["--dataset", "cora", "--dim", "32", "--hidden", "16", "--classes", "7", "--epochs", "3",
                            "--model", "gcn", "--seed", "0"]) == 0
    assert "Train (ms):" in capfd.readouterr().out
    assert main_tcgnn.main(["--dataset", "citeseer", "--dim", "32", "--hidden", "32", "--classes", "6", "--epochs", "2",
                            "--num_layers", "4", "--model", "agnn", "--prep", "gpu", "--seed", "0"]) == 0
    assert "Train (ms):" in capfd.readouterr().out
