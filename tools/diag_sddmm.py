#!/usr/bin/env python3
"""Diagnostic: run-to-run reproducibility / exactness of SDDMM and weighted SpMM on integer data, with the position
(window, tile, group slot) of every mismatching edge.  usage: diag_sddmm.py <workload|rmat:N:E> [D] [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tc-gnn_atc23_b200"))
import torch
import graphgen, TCGNN

name = sys.argv[1] if len(sys.argv) > 1 else "rmat-10m-200m"
dev = torch.device("cuda")
if name.startswith("rmat:") or name.startswith("uniform:"):
    kind, n, nnz = name.split(":")[0], int(name.split(":")[1]), int(name.split(":")[2]); d = 256
else:
    n, nnz, d, kind = graphgen.WORKLOADS[name]
d = int(sys.argv[2]) if len(sys.argv) > 2 else d
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
rp, ci = graphgen.synthetic_graph(n, nnz, kind=kind, seed=0, device=dev)
e = ci.numel()
bp = torch.zeros((n + 15) // 16, dtype=torch.int32, device=dev)
e2c = torch.zeros(e, dtype=torch.int32, device=dev); e2r = torch.zeros(e, dtype=torch.int32, device=dev)
fd = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(fd, 1)
TCGNN.preprocess_gpu(ci, rp, n, 16, 8, bp, e2c, e2r)
os.dup2(saved, 1)
g = (rp, ci, bp, e2c, e2r)
gen = torch.Generator(device=dev).manual_seed(7)
x = torch.randint(-4, 5, (n, d), generator=gen, device=dev).float()
want = torch.empty(e, device=dev)
step = 1 << 21
for s in range(0, e, step):
    want[s:s + step] = (x[e2r[s:s + step].long()] * x[ci[s:s + step].long()]).sum(dim=1)
win_tiles = torch.clamp(bp, min=1).long()
def where(idx):
    r = e2r[idx].long(); w = r // 16; c = e2c[idx].long(); t = c // 8
    return [f"e={int(i)} row={int(a)} win={int(b)} rowinwin={int(a % 16)} tile={int(tt)}/{int(win_tiles[b])} grp={int(tt // 16)} slot={int(tt % 16)} col={int(cc % 8)}"
            for i, a, b, tt, cc in zip(idx, r, w, t, c)]
bad_total = 0
for rep in range(reps):
    ef = TCGNN.forward_ef(x, *g)[0]
    ne = ef != want
    k = int(ne.sum())
    bad_total += k
    print(f"{name} D={d} sddmm run {rep}: {k} of {e} edges wrong", flush=True)
    if k:
        idx = torch.nonzero(ne).flatten()[:24]
        for ln in where(idx): print("   ", ln)
        print("    got", ef[idx[:6]].tolist(), "want", want[idx[:6]].tolist(), flush=True)
w = torch.randint(-3, 4, (e,), generator=gen, device=dev).float()
a = torch.sparse_csr_tensor(rp.long(), ci.long(), w.double(), size=(n, n))
wantY = torch.sparse.mm(a, x[:, :32].double())
for rep in range(reps):
    y = TCGNN.forward_AGNN(x, rp, ci, w.reshape(1, -1), bp, e2c, e2r)[0]
    nr = (y[:, :32].double() != wantY).any(dim=1)
    k = int(nr.sum())
    bad_total += k
    print(f"{name} D={d} weighted spmm run {rep}: {k} of {n} rows wrong (first 32 columns)", flush=True)
    if k:
        rows = torch.nonzero(nr).flatten()[:12]
        print("    rows", rows.tolist(), "windows", (rows // 16).tolist(), "tiles/win", win_tiles[rows // 16].tolist())
for rep in range(2):
    y = TCGNN.forward(x, *g)[0]
    a1 = torch.sparse_csr_tensor(rp.long(), ci.long(), torch.ones(e, dtype=torch.float64, device=dev), size=(n, n))
    k = int((y[:, :32].double() != torch.sparse.mm(a1, x[:, :32].double())).any(dim=1).sum())
    bad_total += k
    print(f"{name} D={d} spmm run {rep}: {k} rows wrong", flush=True)
print("TOTAL_BAD", bad_total)
