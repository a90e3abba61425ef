"""TEST / BENCH INFRASTRUCTURE ONLY (oracle/): CPU restatement of the reference's `dgl_baseline` GCN run, the
"cora GCN 2-layer hidden=16 via dgl_baseline on CPU" plumbing configuration of BASELINE.json (configs[0]).

What it restates (paths relative to /root/reference):
  * dgl_baseline/gcn.py:14-36     GCN = GraphConv(in, hidden, relu) -> [GraphConv(hidden, hidden, relu)] * (L - 2)
                                  -> GraphConv(hidden, classes), every layer allow_zero_in_degree=True
  * dgl_baseline/train.py:57-88   CrossEntropyLoss, Adam(lr=1e-2, weight_decay=5e-4), model.train(), 3 dry forward
                                  passes, then --n-epochs timed steps, prints "Train (ms): ..."
  * dgl_baseline/dataset.py:76-84 features = randn(N, dim), labels = ones(N) (here seeded)

Third-party arithmetic: `dgl.nn.pytorch.GraphConv` (DGL is not installed, not vendored and its version is unpinned:
README.md:49, docker/dockerfile:23 -> PARITY UNPINNED, timing only, as the reference itself only compares times).
Restated from DGL's published definition of GraphConv with norm='both' (its default): h_i = b + sum_{j in N_in(i)}
(d_out(j) d_in(i))^{-1/2} x_j W, degrees clamped to >= 1; when in_feats > out_feats the projection X W is applied
before the aggregation, otherwise after; weight Glorot-uniform, bias zero.  The aggregation is `torch.sparse.mm` on a
CSR adjacency on the host cores -- the same thing DGL's CPU backend ends up calling.

Nothing under tc-gnn_atc23_b200/ imports this file."""
from __future__ import annotations

import os
import time
import warnings

import numpy as np
import torch


class GraphConvCPU(torch.nn.Module):
    def __init__(self, in_feats: int, out_feats: int, activation=None):
        super().__init__()
        self.in_feats, self.out_feats, self.activation = in_feats, out_feats, activation
        self.weight = torch.nn.Parameter(torch.empty(in_feats, out_feats))
        self.bias = torch.nn.Parameter(torch.zeros(out_feats))
        torch.nn.init.xavier_uniform_(self.weight)

    def forward(self, adj_t: torch.Tensor, out_norm: torch.Tensor, in_norm: torch.Tensor, feat: torch.Tensor):
        """adj_t: CSR [N, N] with adj_t[i, j] = 1 for every edge j -> i (rows aggregate their in-neighbours)."""
        h = feat * out_norm
        if self.in_feats > self.out_feats:
            h = torch.sparse.mm(adj_t, h @ self.weight)
        else:
            h = torch.sparse.mm(adj_t, h) @ self.weight
        h = h * in_norm + self.bias
        return self.activation(h) if self.activation is not None else h


class GCNCPU(torch.nn.Module):
    def __init__(self, in_feats, n_hidden, n_classes, n_layers):
        super().__init__()
        self.layers = torch.nn.ModuleList([GraphConvCPU(in_feats, n_hidden, torch.relu)])
        for _ in range(n_layers - 2):
            self.layers.append(GraphConvCPU(n_hidden, n_hidden, torch.relu))
        self.layers.append(GraphConvCPU(n_hidden, n_classes))

    def forward(self, adj_t, out_norm, in_norm, features):
        h = features
        for layer in self.layers:
            h = layer(adj_t, out_norm, in_norm, h)
        return h


def build_graph(src: np.ndarray, dst: np.ndarray, num_nodes: int):
    """DGL keeps multi-edges (dataset.py:68 add_edges): adj_t[dst, src] counts them, degrees count them too."""
    key = dst.astype(np.int64) * num_nodes + src.astype(np.int64)
    uniq, cnt = np.unique(key, return_counts=True)
    rows, cols = uniq // num_nodes, uniq % num_nodes
    crow = np.zeros(num_nodes + 1, dtype=np.int64)
    np.cumsum(np.bincount(rows, minlength=num_nodes), out=crow[1:])
    warnings.filterwarnings("ignore", message=".*[Ss]parse.*")
    adj_t = torch.sparse_csr_tensor(torch.from_numpy(crow), torch.from_numpy(cols), torch.from_numpy(cnt.astype(np.float32)),
                                    size=(num_nodes, num_nodes))
    in_deg = np.bincount(dst, minlength=num_nodes).astype(np.float32)
    out_deg = np.bincount(src, minlength=num_nodes).astype(np.float32)
    in_norm = torch.from_numpy(np.maximum(in_deg, 1.0) ** -0.5).unsqueeze(1)
    out_norm = torch.from_numpy(np.maximum(out_deg, 1.0) ** -0.5).unsqueeze(1)
    return adj_t, out_norm, in_norm


def run(num_nodes=2708, num_edges=10858, dim=1433, n_hidden=16, n_classes=7, n_layers=2, n_epochs=100, seed=0,
        threads=None):
    """The dgl_baseline/train.py loop on a seeded cora-sized uniform graph; returns a dict with the timing the
    reference prints ("Train (ms)"), the aggregation-only throughput and the core count used."""
    if threads:
        torch.set_num_threads(int(threads))
    rng = np.random.default_rng(seed)
    half = num_edges // 2
    s, d = rng.integers(0, num_nodes, half), rng.integers(0, num_nodes, half)
    src, dst = np.concatenate([s, d]), np.concatenate([d, s])            # symmetrised like the planetoid graphs
    adj_t, out_norm, in_norm = build_graph(src, dst, num_nodes)
    g = torch.Generator().manual_seed(seed)
    features = torch.randn(num_nodes, dim, generator=g)
    labels = torch.ones(num_nodes, dtype=torch.long)
    torch.manual_seed(seed)
    model = GCNCPU(dim, n_hidden, n_classes, n_layers)
    loss_fcn = torch.nn.CrossEntropyLoss()
    optimizer = torch.optim.Adam(model.parameters(), lr=1e-2, weight_decay=5e-4)
    model.train()
    for _ in range(3):                                                    # dry run (train.py:69-70)
        model(adj_t, out_norm, in_norm, features)
    t0 = time.perf_counter()
    loss = None
    for _ in range(n_epochs):
        logits = model(adj_t, out_norm, in_norm, features)
        loss = loss_fcn(logits, labels)
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()
    dur = time.perf_counter() - t0
    # aggregation alone (what the GPU path replaces): A^T X at the hidden width
    x = torch.randn(num_nodes, n_hidden, generator=g)
    reps = 50
    t1 = time.perf_counter()
    for _ in range(reps):
        torch.sparse.mm(adj_t, x)
    agg = (time.perf_counter() - t1) / reps
    nnz = int(adj_t.values().numel())
    return {"config": f"cora-like GCN {n_layers}-layer hidden={n_hidden} (N={num_nodes}, {len(src)} edges, in-dim {dim}, "
                      f"{n_classes} classes), restated dgl_baseline/train.py on CPU",
            "train_ms_per_epoch": dur * 1e3 / n_epochs, "epochs": n_epochs, "final_loss": float(loss.detach()),
            "aggregation_edges_per_s": nnz / agg, "aggregation_ms": agg * 1e3, "aggregation_dim": n_hidden,
            "cores": torch.get_num_threads(), "host_cpus": os.cpu_count(), "kind": "port",
            "parity": "unpinned (DGL absent, version unpinned): timing only, like the reference"}


def run_best(n_epochs=100, candidates=None, **kw):
    """torch's CPU kernels can get slower with more threads on a small graph (and on a box whose cores are shared):
    try a few thread counts for a handful of epochs, then run the reference's loop with the fastest and say which."""
    ncpu = os.cpu_count() or 1
    cands = candidates or sorted({1, min(4, ncpu), min(16, ncpu), ncpu})
    saved = torch.get_num_threads()
    try:
        probe = {t: run(n_epochs=5, threads=t, **kw)["train_ms_per_epoch"] for t in cands}
        best = min(probe, key=probe.get)
        out = run(n_epochs=n_epochs, threads=best, **kw)
        out["thread_probe_ms_per_epoch"] = {str(k): round(v, 3) for k, v in probe.items()}
        return out
    finally:
        torch.set_num_threads(saved)


if __name__ == "__main__":
    r = run_best()
    print("Train (ms): {:.3f}".format(r["train_ms_per_epoch"]))
    print(r)
