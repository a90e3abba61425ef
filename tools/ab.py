#!/usr/bin/env python3
"""Device timing of several (op, workload) pairs in one process (A/B runs of library variants, see build_variant.py):
    LD_LIBRARY_PATH=$PWD/variants/<name> python tools/ab.py --tag <name> spmm:reddit-like-rmat sddmm:reddit-like-uniform"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tc-gnn_atc23_b200"))
import torch
import graphgen, TCGNN

ap = argparse.ArgumentParser()
ap.add_argument("items", nargs="+")
ap.add_argument("--iters", type=int, default=7)
ap.add_argument("--tag", default="")
a = ap.parse_args()
dev = torch.device("cuda")
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
graphs = {}
for item in a.items:
    op, wl = item.split(":")
    n, nnz, d, kind = graphgen.WORKLOADS[wl]
    if wl not in graphs:
        graphs.clear(); TCGNN.clear_plan_cache(); torch.cuda.empty_cache()
        rp, ci = graphgen.synthetic_graph(n, nnz, kind=kind, seed=0, device=dev)
        e = ci.numel()
        bp = torch.zeros((n + 15) // 16, dtype=torch.int32, device=dev)
        e2c = torch.zeros(e, dtype=torch.int32, device=dev); e2r = torch.zeros(e, dtype=torch.int32, device=dev)
        fd = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(fd, 1)
        TCGNN.preprocess_gpu(ci, rp, n, 16, 8, bp, e2c, e2r)
        os.dup2(saved, 1)
        graphs[wl] = (rp, ci, bp, e2c, e2r, graphgen.features(n, d, device=dev))
    rp, ci, bp, e2c, e2r, x = graphs[wl]
    g = (rp, ci, bp, e2c, e2r)
    aw = torch.full((1, 1), 0.01, device=dev)
    att = torch.rand(1, ci.numel(), device=dev) if op == "wspmm" else None
    run = {"spmm": lambda: TCGNN.forward(x, *g)[0], "sddmm": lambda: TCGNN.forward_ef(x, *g)[0],
           "agnn": lambda: TCGNN.forward_AGNN_fused(x, rp, ci, aw, bp, e2c, e2r, False)[0],
           "wspmm": lambda: TCGNN.forward_AGNN(x, rp, ci, att, bp, e2c, e2r)[0]}[op]
    run(); torch.cuda.synchronize()
    ts = []
    for _ in range(a.iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print(f"{a.tag:>12s} {op:>6s} {wl:<20s} min_ms={min(ts):.3f} med_ms={sorted(ts)[len(ts)//2]:.3f}", flush=True)
