#!/usr/bin/env python3
"""Stress: many SDDMM runs on integer data, histogram of where wrong edges sit (tile slot in its group, column in the
tile, row in the window).  usage: stress_sddmm.py <workload> <D> <reps>"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tc-gnn_atc23_b200"))
import torch
import graphgen, TCGNN
name, d, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda")
if ":" in name:
    kind, n, nnz = name.split(":")[0], int(name.split(":")[1]), int(name.split(":")[2])
else:
    n, nnz, _, kind = graphgen.WORKLOADS[name]
rp, ci = graphgen.synthetic_graph(n, nnz, kind=kind, seed=0, device=dev)
e = ci.numel()
bp = torch.zeros((n + 15) // 16, dtype=torch.int32, device=dev)
e2c = torch.zeros(e, dtype=torch.int32, device=dev); e2r = torch.zeros(e, dtype=torch.int32, device=dev)
fd = os.open(os.devnull, os.O_WRONLY); saved = os.dup(1); os.dup2(fd, 1)
TCGNN.preprocess_gpu(ci, rp, n, 16, 8, bp, e2c, e2r)
os.dup2(saved, 1)
g = (rp, ci, bp, e2c, e2r)
x = torch.randint(-4, 5, (n, d), generator=torch.Generator(device=dev).manual_seed(7), device=dev).float()
want = torch.empty(e, device=dev)
step = 1 << 21
for s in range(0, e, step):
    want[s:s + step] = (x[e2r[s:s + step].long()] * x[ci[s:s + step].long()]).sum(dim=1)
hist = collections.Counter(); bad_runs = 0; total = 0
op = sys.argv[4] if len(sys.argv) > 4 else "sddmm"
if op != "sddmm":
    gen = torch.Generator(device=dev).manual_seed(9)
    w = torch.randint(-3, 4, (e,), generator=gen, device=dev).float()
    a = torch.sparse_csr_tensor(rp.long(), ci.long(), (w if op == "wspmm" else torch.ones_like(w)).double(), size=(n, n))
    wantY = torch.sparse.mm(a, x[:, :32].double())
    att = w.reshape(1, -1)
    bad_rows = collections.Counter()
    for rep in range(reps):
        y = TCGNN.forward_AGNN(x, rp, ci, att, bp, e2c, e2r)[0] if op == "wspmm" else TCGNN.forward(x, *g)[0]
        nr = torch.nonzero((y[:, :32].double() != wantY).any(dim=1)).flatten()
        if nr.numel():
            bad_runs += 1; total += nr.numel()
            print(f"  run {rep}: {nr.numel()} wrong rows {nr[:8].tolist()} windows {(nr[:8] // 16).tolist()}", flush=True)
    print(f"{name} D={d} op={op} team={os.environ.get('TCGNN_SPMM_TEAM','auto')} lag={os.environ.get('TCGNN_SPMM_LAG','auto')}: "
          f"{bad_runs} of {reps} runs wrong, {total} rows", flush=True)
    sys.exit(0)
for rep in range(reps):
    ef = TCGNN.forward_ef(x, *g)[0]
    idx = torch.nonzero(ef != want).flatten()
    if idx.numel():
        bad_runs += 1; total += idx.numel()
        t = e2c[idx].long() // 8
        for sl, c, r in zip((t % 16).tolist(), (e2c[idx].long() % 8).tolist(), (e2r[idx].long() % 16).tolist()):
            hist[(sl, c)] += 1
        tiles = sorted({(int(w), int(tt)) for w, tt in zip((e2r[idx].long() // 16).tolist(), t.tolist())})
        print(f"  run {rep}: {idx.numel()} wrong edges in tiles (window, tile) {tiles[:6]}", flush=True)
tag = f"team={os.environ.get('TCGNN_SDDMM_TEAM','2')} dbg={os.environ.get('TCGNN_SDDMM_DBG','0')}"
print(f"{name} D={d} {tag}: {bad_runs} of {reps} runs wrong, {total} edges; (slot, col) histogram {sorted(hist.items())}", flush=True)
