#!/usr/bin/env python3
"""Where does a sharded step's time go?  (torchrun, N GPUs.)  Each variant is captured as a CUDA graph and replayed;
times are CUDA events around the replay, mean over 20 steps after a barrier, max over ranks.
  push_ce1      round + copy-engine pushes to all peers on ONE stream + flags + waits for all sources (no products)
  push_ceN      same, one copy stream per peer
  push_sm       round fused with SM stores into every peer (round_tf32_into per peer, rotated) + flags + waits
  products_1x   own + one product per source on parallel streams, data already resident (no exchange, no waits)
  products_grp  own + default groups, data resident
  products_one  the legacy single panel kernel on a gathered matrix
  step          the shipped overlapped step
"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tc-gnn_atc23_b200"))
import torch, torch.distributed as dist
import graphgen, TCGNN
from sharding import RowPanel

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"])); torch.cuda.set_device(dev)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
dist.init_process_group("nccl", device_id=dev)
wl = sys.argv[1] if len(sys.argv) > 1 else "reddit-like-rmat"
n, nnz, d, kind = graphgen.WORKLOADS[wl]
rp, ci = graphgen.synthetic_graph(n, nnz, kind=kind, seed=0, device=dev)
os.environ["TCGNN_EXCHANGE_GROUPS"] = ",".join(["1"] * (world - 1))
panel = RowPanel(rp, ci, rank, world, device=dev)
x = graphgen.features(n, d, seed=0, device=dev)[panel.row_base:panel.row_base + panel.num_rows].contiguous()
for _ in range(6):
    y = panel.aggregate(x)                       # sets everything up, captures the shipped step
torch.cuda.synchronize(); dist.barrier()
st = panel._ovl[d]; subs = panel._sub; w, me = world, rank
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
copy_streams = [torch.cuda.Stream(device=dev) for _ in range(w - 1)]


def pushes(b, mode):
    cur = torch.cuda.current_stream(dev)
    xr = st.xr[b]
    st.step_dev.add_(1)
    if mode != "sm":
        TCGNN.round_tf32_into(x, xr.data_ptr(), d, False)
    for k in range(1, w):
        q = (me + k) % w
        cs = st.copy_stream if mode in ("ce1", "sm") else copy_streams[k - 1]
        cs.wait_stream(cur)
        with torch.cuda.stream(cs):
            nq = int(st.need[q, me]); o = int(st.offs[q, me])
            if mode == "sm":
                TCGNN.round_tf32_into(x, st.peer_recv[q][b, o:o + nq].data_ptr(), d, False)
            else:
                st.peer_recv[q][b, o:o + nq].copy_(xr, non_blocking=True)
            st.peer_flags[q][me:me + 1].copy_(st.step_dev, non_blocking=True)
    for k in range(1, w):
        TCGNN.stream_wait_flag_dev(st.flags, (me - k) % w, st.step_dev, 20000, st.err)
    for cs in ([st.copy_stream] if mode in ("ce1", "sm") else copy_streams):
        cur.wait_stream(cs)


def products(b, groups):
    cur = torch.cuda.current_stream(dev)
    yy = torch.zeros((panel.num_rows, d), device=dev)
    TCGNN.source_forward(st.xr[b], *subs[me]["graph"], x_is_tf32=True, accumulate_into=yy)
    for s_ in st.product_streams:
        s_.wait_stream(cur)
    for gi, g in enumerate(groups):
        with torch.cuda.stream(st.product_streams[gi % len(st.product_streams)]):
            o = int(st.offs[me, g["sources"][0]])
            TCGNN.source_forward(st.recv[b, o:o + g["ncols"]], *g["graph"], x_is_tf32=True, accumulate_into=yy)
    for s_ in st.product_streams:
        cur.wait_stream(s_)
    return yy


def bench(name, fn, reps=20):
    fn(0); fn(1); torch.cuda.synchronize(); dist.barrier()
    graphs = []
    for b in (0, 1):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            fn(b)
        graphs.append(g)
    for i in range(4):
        graphs[i & 1].replay()
    torch.cuda.synchronize(); dist.barrier()
    tot = 0.0
    for i in range(reps):
        flush.zero_(); torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graphs[i & 1].replay(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    t = torch.tensor([tot / reps], device=dev, dtype=torch.float64)
    tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tmin = t.clone(); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"{name:14s} max over ranks {float(tmax):.3f} ms   min over ranks {float(tmin):.3f} ms", flush=True)


singles = st.groups
os.environ.pop("TCGNN_EXCHANGE_GROUPS")
panel._groups = None
grouped = panel.build_group_subgraphs()
xa = panel.all_gather(x, round_tf32=True)
rows = torch.tensor([panel.num_rows], device=dev); allrows = [torch.zeros_like(rows) for _ in range(w)]; dist.all_gather(allrows, rows)
if rank == 0:
    print(f"{wl} N={w} panel rows {[int(r) for r in allrows]} groups {[g['sources'] for g in grouped]}", flush=True)
bench("push_ce1", lambda b: pushes(b, "ce1"))
bench("push_ceN", lambda b: pushes(b, "ceN"))
bench("push_sm", lambda b: pushes(b, "sm"))
bench("products_1x", lambda b: products(b, singles))
bench("products_grp", lambda b: products(b, grouped))
bench("products_one", lambda b: panel.spmm(xa, x_is_tf32=True))
bench("step", lambda b: panel._overlap_step(x, st, b))
dist.barrier(); dist.destroy_process_group()
