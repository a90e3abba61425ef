#!/usr/bin/env python3
"""Phase timing of the sharded step (torchrun, >= 2 GPUs): barrier / round+push / barrier / panel SpMM."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tc-gnn_atc23_b200"))
import torch, torch.distributed as dist
import graphgen, TCGNN
from sharding import RowPanel
import torch.distributed._symmetric_memory as symm_mem

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"])); torch.cuda.set_device(dev)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=dev)
wl = sys.argv[1] if len(sys.argv) > 1 else "reddit-like-rmat"
n, nnz, d, kind = graphgen.WORKLOADS[wl]
rp, ci = graphgen.synthetic_graph(n, nnz, kind=kind, seed=0, device=dev)
panel = RowPanel(rp, ci, rank, world, device=dev)
x = graphgen.features(n, d, seed=0, device=dev)[panel.row_base:panel.row_base + panel.num_rows].contiguous()
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for mode in ("nccl", "p2p", "p2p2"):
    os.environ["TCGNN_EXCHANGE"] = mode
    for _ in range(3):
        panel.spmm(panel.all_gather(x, round_tf32=True), x_is_tf32=True)
    torch.cuda.synchronize(); dist.barrier()
    ev = lambda: torch.cuda.Event(enable_timing=True)
    tot = [0.0, 0.0, 0.0]; host = 0.0
    for it in range(10):
        flush.zero_(); torch.cuda.synchronize(); dist.barrier()
        e = [ev() for _ in range(3)]
        t0 = time.perf_counter()
        e[0].record(); xa = panel.all_gather(x, round_tf32=True); e[1].record()
        y = panel.spmm(xa, x_is_tf32=True); e[2].record()
        host += time.perf_counter() - t0
        torch.cuda.synchronize()
        tot[0] += e[0].elapsed_time(e[1]); tot[1] += e[1].elapsed_time(e[2]); tot[2] += e[0].elapsed_time(e[2])
    msg = f"[{mode}] rank {rank} rows {panel.num_rows}: exchange {tot[0]/10:.3f} ms  spmm {tot[1]/10:.3f} ms  step {tot[2]/10:.3f} ms  host-issue {host/10*1e3:.3f} ms"
    for r in range(world):
        if r == rank: print(msg, flush=True)
        dist.barrier()
dist.destroy_process_group()
