#!/usr/bin/env python3
"""Quick device timing of one operator on a synthetic workload (iteration tool, not the benchmark)."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tc-gnn_atc23_b200"))
import torch
import graphgen, TCGNN

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="reddit-like-uniform")
ap.add_argument("--op", default="spmm")
ap.add_argument("--dim", type=int, default=0)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--tag", default="")
a = ap.parse_args()
n, nnz, d, kind = graphgen.WORKLOADS[a.workload]
d = a.dim or d
dev = torch.device("cuda")
rp, ci = graphgen.synthetic_graph(n, nnz, kind=kind, seed=0, device=dev)
e = ci.numel()
bp = torch.zeros((n + 15) // 16, dtype=torch.int32, device=dev)
e2c = torch.zeros(e, dtype=torch.int32, device=dev); e2r = torch.zeros(e, dtype=torch.int32, device=dev)
fd = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(fd, 1)
TCGNN.preprocess_gpu(ci, rp, n, 16, 8, bp, e2c, e2r)
os.dup2(saved, 1)
g = (rp, ci, bp, e2c, e2r)
x = graphgen.features(n, d, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
att = None
aw = torch.full((1, 1), 0.01, device=dev)
att_tile = None
x_host = y_host = None
def run():
    if a.op == "spmm": return TCGNN.forward(x, *g)[0]
    if a.op == "sddmm": return TCGNN.forward_ef(x, *g)[0]
    if a.op == "agnn": return TCGNN.forward_AGNN_fused(x, rp, ci, aw, bp, e2c, e2r, False)[0]       # fused entry
    if a.op == "agnn3":                                                                              # the three calls
        ef = TCGNN.forward_ef(x, *g)[0]
        return TCGNN.forward_AGNN(x, rp, ci, torch.mm(ef.unsqueeze(-1), aw).transpose(0, 1).contiguous(), bp, e2c, e2r)[0]
    if a.op == "wspmm_tile": return TCGNN.forward_AGNN_tile(x, rp, ci, att_tile, bp, e2c, e2r)[0]   # tile-ordered weights
    if a.op == "spmm_host": return TCGNN.forward_host(x_host, *g, y_host=y_host, sync=False)
    if a.op == "spmm_T": return TCGNN.backward_T(x, *g)[0]
    return TCGNN.forward_AGNN(x, rp, ci, att, bp, e2c, e2r)[0]
if a.op == "wspmm": att = torch.rand(1, e, device=dev)
if a.op == "wspmm_tile": att_tile = TCGNN.forward_AGNN_fused(x, rp, ci, aw, bp, e2c, e2r, False)[1]
if a.op == "spmm_host":
    x_host = x.cpu().pin_memory(); y_host = torch.empty(n, d).pin_memory()
run(); torch.cuda.synchronize()
ts = []
for _ in range(a.iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
tiles = TCGNN.plan_info(*g)[3]
ms = min(ts)
print(f"{a.tag or os.environ.get('TCGNN_ABLATE','0'):>10s} {a.workload} op={a.op} D={d} E={e} tiles={tiles} min_ms={ms:.3f} med_ms={sorted(ts)[len(ts)//2]:.3f} "
      f"ns/tile/SM={ms*1e6/tiles*148:.1f} Gedges/s={e/ms/1e6:.2f}", flush=True)
