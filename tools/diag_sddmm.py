#!/usr/bin/env python3
"""Diagnostic: is SDDMM bit-reproducible run to run, and fused == separate, on a named workload with normal data?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tc-gnn_atc23_b200"))
import torch
import graphgen, TCGNN

name = sys.argv[1] if len(sys.argv) > 1 else "rmat-10m-200m"
n, nnz, d, kind = graphgen.WORKLOADS[name]
dev = torch.device("cuda")
rp, ci = graphgen.synthetic_graph(n, nnz, kind=kind, seed=0, device=dev)
e = ci.numel()
bp = torch.zeros((n + 15) // 16, dtype=torch.int32, device=dev)
e2c = torch.zeros(e, dtype=torch.int32, device=dev); e2r = torch.zeros(e, dtype=torch.int32, device=dev)
fd = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(fd, 1)
TCGNN.preprocess_gpu(ci, rp, n, 16, 8, bp, e2c, e2r)
os.dup2(saved, 1)
g = (rp, ci, bp, e2c, e2r)
x = torch.randn(n, d, generator=torch.Generator(device=dev).manual_seed(8), device=dev) * 0.1
aw = torch.full((1, 1), 0.37, device=dev)
ef1 = TCGNN.forward_ef(x, *g)[0]
ef2 = TCGNN.forward_ef(x, *g)[0]
y_f, att, ef_f = TCGNN.forward_AGNN_fused(x, rp, ci, aw, bp, e2c, e2r, True)
ef3 = TCGNN.forward_ef(x, *g)[0]
torch.cuda.synchronize()
def cmp(a, b, what):
    ne = a != b
    k = int(ne.sum())
    msg = f"{what}: {k} of {a.numel()} differ"
    if k:
        idx = torch.nonzero(ne).flatten()
        i0 = int(idx[0]); i1 = int(idx[-1])
        md = float((a - b).abs().max())
        rows = e2r[idx[:5]].tolist()
        msg += f"; max|diff| {md:.3e}; first edge {i0} last {i1}; rows of first five {rows}; a={a[idx[:3]].tolist()} b={b[idx[:3]].tolist()}"
        msg += f"; nan a {int(torch.isnan(a).sum())} b {int(torch.isnan(b).sum())}"
    print(msg, flush=True)
cmp(ef1, ef2, f"{name} forward_ef run1 vs run2")
cmp(ef1, ef_f, f"{name} forward_ef vs fused edge_feature")
cmp(ef1, ef3, f"{name} forward_ef run1 vs run3 (after fused)")
info = TCGNN.plan_info(*g)
print("pairs", info[5], "edges", e, "tiles", info[3])
