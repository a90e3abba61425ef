#!/usr/bin/env python3
"""TC blocks and SpMM time before / after node reordering on the named workloads (GPU), in the format of the
reference's logs/16x8_reduction.csv (dataset, origin = plain 16x8 tiling, reduced = SGT) plus the reordered columns.

    python tools/reorder_report.py [--out gpurun_out/reorder_tc_blocks.csv] [workload ...]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tc-gnn_atc23_b200"))
import torch
import graphgen, reorder, TCGNN

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "reorder_tc_blocks.csv"))
ap.add_argument("workloads", nargs="*", default=["reddit-like-rmat", "reddit-like-uniform", "products-like-rmat", "rmat-10m-200m"])
a = ap.parse_args()
dev = torch.device("cuda")


def spmm_ms(rp, ci, n, d):
    e = ci.numel()
    bp = torch.zeros((n + 15) // 16, dtype=torch.int32, device=dev)
    e2c = torch.zeros(e, dtype=torch.int32, device=dev); e2r = torch.zeros(e, dtype=torch.int32, device=dev)
    fd = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(fd, 1)
    TCGNN.preprocess_gpu(ci, rp, n, 16, 8, bp, e2c, e2r)
    os.dup2(saved, 1); os.close(fd); os.close(saved)
    g = (rp, ci, bp, e2c, e2r)
    x = graphgen.features(n, d, device=dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    TCGNN.forward(x, *g); torch.cuda.synchronize()
    ts = []
    for _ in range(3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); TCGNN.forward(x, *g); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    tiles = TCGNN.plan_info(*g)[3]
    TCGNN.clear_plan_cache()
    return min(ts), tiles


lines = ["dataset,nnz,origin,reduced,reduction (%),nnz_per_block,spmm_ms,method,reordered,reordered_reduction (%),"
         "reordered_nnz_per_block,reordered_spmm_ms,reorder_s"]
for name in a.workloads:
    n, nnz, d, kind = graphgen.WORKLOADS[name]
    rp, ci = graphgen.synthetic_graph(n, nnz, kind=kind, seed=0, device=dev)
    origin = reorder.naive_tc_blocks(rp, ci)
    base = reorder.count_tc_blocks(rp, ci)
    ms0, tiles0 = spmm_ms(rp, ci, n, d)
    assert tiles0 == base, (tiles0, base)
    for method in ("hub", "minhash", "degree"):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        rp2, ci2, perm, rep = reorder.reorder_graph(rp, ci, method)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        ms1, tiles1 = spmm_ms(rp2, ci2, n, d)
        assert tiles1 == rep["tc_blocks_after"]
        lines.append(f"{name},{ci.numel()},{origin},{base},{100 * (origin - base) / origin:.2f},{ci.numel() / base:.2f},{ms0:.3f},"
                     f"{method},{tiles1},{100 * (base - tiles1) / base:.2f},{ci.numel() / tiles1:.2f},{ms1:.3f},{dt:.2f}")
        print(lines[-1], flush=True)
        del rp2, ci2, perm
    del rp, ci
    torch.cuda.empty_cache()
os.makedirs(os.path.dirname(a.out), exist_ok=True)
with open(a.out, "w") as fh:
    fh.write("\n".join(lines) + "\n")
