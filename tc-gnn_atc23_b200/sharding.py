"""1-D row-panel (destination-node) sharding of the aggregation path across the GPUs of one box.

New work -- the reference is single process / single GPU.  Design (DESIGN.md "multi-GPU"):

  * rank g owns the rows [bounds[g], bounds[g+1]) of the graph; boundaries fall on multiples of
    BLK_H = 16 rows, so every 16-row window belongs to exactly one rank and a panel's SGT arrays are
    exactly the panel's slice of the whole graph's SGT arrays (window-local ranks, per-window tile
    counts) -- sharded results are bit-identical to the single-GPU ones;
  * boundaries are chosen on the prefix sum of the CSR row pointer so the stored non-zeros (= the
    gather work) per rank are balanced, not the row counts (R-MAT rows are heavily skewed);
  * column ids stay global.  Per layer the only exchange is ONE all-gather of the layer input
    (every rank contributes its panel of X, receives the others') over NVLink; the output of
    SpMM stays sharded -- it is the rank's panel of the next layer's input;
  * GCN/GIN/SAG aggregation (`aggregate`) overlaps that exchange with the kernels: the panel's sub-graph
    is split once by SOURCE panel, Y = sum_p A[panel, cols of p] . X[p]; the own-panel product starts
    at once, every other product as soon as that source's rows have landed (copy-engine pushes into
    symmetric memory + a flag, a stream-ordered wait in front of each launch, TCGNN_ACCUMULATE).  A
    source ships only the rows the destination's panel references when that is a small part of its
    panel (packed, column ids remapped) -- the R-MAT 10 M graph references ~1/3 of the rows;
  * AGNN's SDDMM + weighted SpMM take the gathered matrix plus the panel plan
    (`tcgnn_plan_create_panel`) and share one gather.

Host-side logic here is device-agnostic (the world_size-2 gloo tests run it on CPU); the compute
calls go to the `TCGNN` extension and need a GPU -- there is no CPU fallback.
"""
from __future__ import annotations

import os
import sys
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from config import BLK_H, BLK_W


def partition_rows(row_ptr: torch.Tensor, world_size: int, blk_h: int = BLK_H,
                   block_partition: Optional[torch.Tensor] = None, window_cost: int = 3,
                   send_cost_per_row: float = 0.0) -> List[int]:
    """Row boundaries [b_0 = 0, ..., b_world = N], each a multiple of `blk_h` (except N).  Deterministic,
    identical on every rank.

    Without `block_partition` the stored non-zeros per panel are balanced.  With the graph's SGT tile counts
    (`blockPartition`, one entry per `blk_h`-row window) the panels are balanced on what the kernels' time is
    actually proportional to: TC blocks + `window_cost` per window (the same cost model the CTA slices inside a
    GPU use, plan.cu) -- on R-MAT graphs the hub panel has 2-3x more non-zeros per TC block than the tail
    panel, and non-zero balancing left the slowest rank 37 % behind on 2 GPUs."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    n = int(row_ptr.numel()) - 1
    last = n // blk_h * blk_h if n >= blk_h else 0
    if block_partition is not None:
        cost = torch.clamp(block_partition.to(torch.int64).cpu(), min=1) + int(window_cost)
        if send_cost_per_row > 0 and world_size > 1:
            return _partition_minmax(cost.numpy(), window_send_cost(row_ptr, world_size, send_cost_per_row, blk_h),
                                     world_size, n, blk_h)
        pre = torch.cumsum(cost, 0)                      # cost of windows [0, w]
        total = int(pre[-1]) if pre.numel() else 0
        bounds = [0]
        for k in range(1, world_size):
            target = (total * k) // world_size
            w = int(torch.searchsorted(pre, torch.tensor(target, dtype=torch.int64), right=False)) + 1
            r = min(max(w * blk_h, bounds[-1]), last)
            bounds.append(r)
        bounds.append(n)
        return bounds
    rp = row_ptr.to(torch.int64).cpu()
    nnz = int(rp[-1] - rp[0])
    bounds = [0]
    for k in range(1, world_size):
        target = int(rp[0]) + (nnz * k) // world_size
        r = int(torch.searchsorted(rp, torch.tensor(target, dtype=torch.int64), right=False))
        r = min(max(r, 0), n)
        r = (r + blk_h // 2) // blk_h * blk_h          # nearest window boundary
        r = min(max(r, bounds[-1]), last)
        bounds.append(r)
    bounds.append(n)
    return bounds


def window_send_cost(row_ptr: torch.Tensor, world_size: int, send_cost_per_row: float, blk_h: int = BLK_H):
    """Send cost of every `blk_h`-row window: a row is shipped to the panels that reference it -- at most
    min(degree, world - 1) of them on a symmetric graph -- and `send_cost_per_row` is the cost of shipping one row to
    ALL world - 1 peers."""
    import numpy as np
    rp = row_ptr.to(torch.int64).cpu().numpy()
    n = len(rp) - 1
    deg = np.minimum(np.diff(rp), world_size - 1).astype(np.float64) / max(world_size - 1, 1)
    nwin = (n + blk_h - 1) // blk_h
    pad = np.zeros(nwin * blk_h, dtype=np.float64)
    pad[:n] = deg
    return pad.reshape(nwin, blk_h).sum(axis=1) * float(send_cost_per_row)


def _partition_minmax(compute, send, world: int, n: int, blk_h: int) -> List[int]:
    """Window cuts minimising max over panels of max(sum compute, sum send): bisection on the bound T, greedy
    feasibility (extend a panel while both sums stay <= T)."""
    import numpy as np
    nwin = len(compute)
    pre_c = np.concatenate([[0.0], np.cumsum(compute, dtype=np.float64)])
    pre_s = np.concatenate([[0.0], np.cumsum(send, dtype=np.float64)])

    def cuts(t):
        out, w = [0], 0
        for _ in range(world):
            # furthest end with compute <= t and send <= t
            hi_c = int(np.searchsorted(pre_c, pre_c[w] + t, side="right")) - 1
            hi_s = int(np.searchsorted(pre_s, pre_s[w] + t, side="right")) - 1
            e = max(min(hi_c, hi_s, nwin), w)
            out.append(e)
            w = e
        return out

    lo = max(pre_c[-1] / world, pre_s[-1] / world)
    hi = max(pre_c[-1], pre_s[-1]) + 1.0
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if cuts(mid)[-1] >= nwin:
            hi = mid
        else:
            lo = mid
    c = cuts(hi)
    c[-1] = nwin
    bounds = [min(w * blk_h, n) for w in c]
    last = n // blk_h * blk_h if n >= blk_h else 0
    return [0] + [min(b, last) for b in bounds[1:-1]] + [n]


def default_send_cost(world_size: int) -> float:
    """Send cost of one feature row in TC-block units: a row of D floats goes to world - 1 peers over NVLink
    (~0.6 TB/s out of a GPU) while a TC block gathers 8 such rows through L2 at ~10 TB/s -- D cancels."""
    v = os.environ.get("TCGNN_SEND_COST")
    if v is not None:
        return float(v)
    return (world_size - 1) / 8.0 * 14.0 if world_size > 2 else 0.0


def _all_ranks_ok(ok: bool, device, group) -> bool:
    """Collective AND of a per-rank success flag (decides fused-versus-fallback identically on every rank)."""
    t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return bool(int(t.item()) == 1)


def calibrated_panel(row_ptr: torch.Tensor, col_idx: torch.Tensor, rank: int, world_size: int, dim: int, device,
                     group=None) -> "RowPanel":
    """Row panel whose boundaries come from MEASURED kernel times instead of an assumed cost per window.

    The panels are first cut with the default model (TC blocks + 3 per window, bounded send volume); every rank times
    the panel SpMM on its cut (a few milliseconds), the ranks exchange (TC blocks, windows, time), fit
    time = a * blocks + b * windows by least squares, and cut again with window cost b / a.  Round 1's fixed cost
    left the slowest of 8 panels 1.33x behind the fastest on the reddit-sized R-MAT graph (tail panels hold 10x the
    windows of the hub panel).  Deterministic across ranks: everybody fits the same gathered numbers."""
    import numpy as np
    import TCGNN
    n = int(row_ptr.numel()) - 1
    nwin_all = (n + BLK_H - 1) // BLK_H
    g_bp = torch.zeros(nwin_all, dtype=torch.int32, device=row_ptr.device)
    g_e2c = torch.zeros(col_idx.numel(), dtype=torch.int32, device=row_ptr.device)
    g_e2r = torch.zeros(col_idx.numel(), dtype=torch.int32, device=row_ptr.device)
    TCGNN.preprocess_panel(col_idx.contiguous(), row_ptr.contiguous(), n, n, BLK_H, BLK_W, g_bp, g_e2c, g_e2r)
    sgt = (g_bp, g_e2c, g_e2r)
    p0 = RowPanel(row_ptr, col_idx, rank, world_size, device=device, sgt=sgt)
    if world_size < 2 or os.environ.get("TCGNN_CALIBRATE", "1") == "0" or not row_ptr.is_cuda:
        return p0
    x = torch.randn(n, dim, device=device)
    xr = TCGNN.round_tf32(x) if dim % 4 == 0 else x
    for _ in range(2):
        p0.spmm(xr, x_is_tf32=dim % 4 == 0)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(device)
    ev0.record()
    for _ in range(4):
        p0.spmm(xr, x_is_tf32=dim % 4 == 0)
    ev1.record()
    torch.cuda.synchronize(device)
    w0 = p0.row_base // BLK_H
    nw = (p0.num_rows + BLK_H - 1) // BLK_H
    mine = torch.tensor([float(torch.clamp(g_bp[w0:w0 + nw], min=1).sum()), float(nw), ev0.elapsed_time(ev1) / 4],
                        dtype=torch.float64, device=device)
    allv = [torch.zeros_like(mine) for _ in range(world_size)]
    dist.all_gather(allv, mine, group=group)
    m = torch.stack(allv).cpu().numpy()
    a_mat = m[:, :2]
    coef, *_ = np.linalg.lstsq(a_mat, m[:, 2], rcond=None)
    a_, b_ = float(coef[0]), float(coef[1])
    # (scaling the window cost by the number of products of the overlapped step -- each product walks all windows --
    # was measured and is worse: 0.71 vs 0.605 ms at 8 GPUs, profiles/r02h_*calibrated*)
    n_products = int(os.environ.get("TCGNN_CALIBRATE_PRODUCTS", "1"))
    wc = 3 if a_ <= 0 else int(round(min(max(b_ / a_ * n_products, 0.0), 512.0)))
    import TCGNN as _T
    _T.clear_plan_cache()
    del x, xr
    if wc == 3:
        return p0
    bounds = partition_rows(row_ptr, world_size, block_partition=g_bp, window_cost=wc,
                            send_cost_per_row=default_send_cost(world_size))
    p1 = RowPanel(row_ptr, col_idx, rank, world_size, bounds=bounds, device=device, sgt=sgt)
    p1.calibration = {"window_cost": wc, "products_per_step": n_products, "fit_ms_per_block": a_, "fit_ms_per_window": b_,
                      "times_before_ms": [round(float(v), 4) for v in m[:, 2]]}
    return p1


class OverlapState:
    """Buffers, step counter and captured graphs of the overlapped exchange for one feature width
    (see RowPanel.aggregate_overlapped)."""


class RowPanel:
    """One rank's row panel of a graph: its CSR slice, its SGT arrays and its exchange plan."""

    def __init__(self, row_ptr: torch.Tensor, col_idx: torch.Tensor, rank: int, world_size: int,
                 bounds: Optional[Sequence[int]] = None, device=None, sgt: Optional[tuple] = None):
        """`row_ptr` / `col_idx`: the whole graph's int32 CSR (any device).  `sgt`: optionally the whole
        graph's (blockPartition, edgeToColumn, edgeToRow) to slice instead of recomputing."""
        self.rank, self.world_size = int(rank), int(world_size)
        self.num_cols = int(row_ptr.numel()) - 1
        if bounds is None and sgt is None and world_size > 1 and row_ptr.is_cuda and self.num_cols > 0:
            # SGT of the whole graph on this GPU (milliseconds), so the panels can be balanced on TC blocks;
            # the panel's arrays are then slices of it (bit-identical to a per-panel SGT)
            import TCGNN
            nwin_all = (self.num_cols + BLK_H - 1) // BLK_H
            g_bp = torch.zeros(nwin_all, dtype=torch.int32, device=row_ptr.device)
            g_e2c = torch.zeros(col_idx.numel(), dtype=torch.int32, device=row_ptr.device)
            g_e2r = torch.zeros(col_idx.numel(), dtype=torch.int32, device=row_ptr.device)
            TCGNN.preprocess_panel(col_idx.contiguous(), row_ptr.contiguous(), self.num_cols, self.num_cols, BLK_H,
                                   BLK_W, g_bp, g_e2c, g_e2r)
            sgt = (g_bp, g_e2c, g_e2r)
        if bounds is not None:
            self.bounds = list(bounds)
        else:
            self.bounds = partition_rows(row_ptr, world_size, block_partition=sgt[0] if sgt is not None else None,
                                         send_cost_per_row=default_send_cost(world_size) if sgt is not None else 0.0)
        if len(self.bounds) != world_size + 1 or self.bounds[0] != 0 or self.bounds[-1] != self.num_cols:
            raise ValueError("bounds must be [0, ..., num_nodes] with world_size + 1 entries")
        for b in self.bounds[1:-1]:
            if b % BLK_H != 0:
                raise ValueError("panel boundaries must be multiples of BLK_H")
        self.row_base = self.bounds[rank]
        self.num_rows = self.bounds[rank + 1] - self.row_base
        device = torch.device(device) if device is not None else row_ptr.device
        self.device = device
        r0, r1 = self.row_base, self.row_base + self.num_rows
        e0, e1 = int(row_ptr[r0]), int(row_ptr[r1])
        self.edge_begin, self.edge_end = e0, e1
        self.row_pointers = (row_ptr[r0:r1 + 1] - row_ptr[r0]).to(torch.int32).to(device).contiguous()
        self.column_index = col_idx[e0:e1].to(torch.int32).to(device).contiguous()
        self.num_edges = e1 - e0
        nwin = (self.num_rows + BLK_H - 1) // BLK_H
        if sgt is not None:
            bp, e2c, e2r = sgt
            w0 = r0 // BLK_H
            self.blockPartition = bp[w0:w0 + nwin].to(device).contiguous()
            self.edgeToColumn = e2c[e0:e1].to(device).contiguous()
            self.edgeToRow = (e2r[e0:e1] - r0).to(torch.int32).to(device).contiguous()
        else:
            import TCGNN   # the SGT is product code (host threads or device), not a fallback
            self.blockPartition = torch.zeros(nwin, dtype=torch.int32, device=device)
            self.edgeToColumn = torch.zeros(self.num_edges, dtype=torch.int32, device=device)
            self.edgeToRow = torch.zeros(self.num_edges, dtype=torch.int32, device=device)
            if self.num_rows > 0:
                TCGNN.preprocess_panel(self.column_index, self.row_pointers, self.num_rows, self.num_cols, BLK_H,
                                       BLK_W, self.blockPartition, self.edgeToColumn, self.edgeToRow)
        self._x_all = None
        self._symm = {}    # feature width -> (symmetric buffer, handle, peer views, multicast?)
        self._symm_unavailable = False
        self._sub = None   # per-source-panel sub-graphs (overlapped exchange)
        self._groups = None  # (group sizes, merged sub-graphs of consecutive arrivals)
        self._ovl = {}     # feature width -> OverlapState
        self._ovl_unavailable = False

    # ------------------------------------------------------------------ exchange
    @property
    def graph(self):
        return (self.row_pointers, self.column_index, self.blockPartition, self.edgeToColumn, self.edgeToRow)

    def panel_rows(self, g: int) -> int:
        return self.bounds[g + 1] - self.bounds[g]

    def all_gather(self, x_local: torch.Tensor, group=None, out: Optional[torch.Tensor] = None,
                   round_tf32: bool = False) -> torch.Tensor:
        """The per-layer exchange: every rank contributes its [num_rows, D] panel, all receive the
        [num_cols, D] matrix (one all-gather; uneven panels).  With `round_tf32` (CUDA, D % 4 == 0) the
        panel is rounded to TF32 before it is sent, so no rank has to round the whole gathered matrix
        again: pass the result to spmm / sddmm with x_is_tf32=True.

        On NCCL groups with `round_tf32` the exchange is FUSED with the rounding pass (env
        TCGNN_EXCHANGE=multicast|p2p2|p2p|nccl, default: multicast when the group supports it, else the
        two-phase balanced push p2p2): the
        gathered matrix lives in symmetric memory, and the kernel that rounds the local panel writes its
        result straight into every GPU's copy -- one multimem store through the NVSwitch, or one pass per
        peer over P2P-mapped memory -- bracketed by two device-side barriers.  No staging copy, no
        broadcast tree: R-MAT panels are balanced on non-zeros, so their row counts (= bytes sent) differ
        2-3x and NCCL's uneven all-gather (grouped broadcasts) took ~0.9 ms for 119 MB on 8 GPUs."""
        if x_local.shape[0] != self.num_rows:
            raise ValueError(f"x_local has {x_local.shape[0]} rows, the panel has {self.num_rows}")
        d = x_local.shape[1]
        mode = os.environ.get("TCGNN_EXCHANGE", "auto")
        if mode in ("overlap",):
            mode = "auto"          # the overlapped path is `aggregate`; a plain gather uses the fused exchange
        if (round_tf32 and out is None and self.world_size > 1 and x_local.is_cuda and d % 4 == 0
                and mode != "nccl" and not self._symm_unavailable and dist.get_backend(group) == "nccl"):
            # Setup (symmetric allocation + rendezvous) can fail on one rank only (out of memory, no P2P access):
            # the ranks agree on fused-versus-NCCL collectively, once per feature width, BEFORE any barrier or data
            # movement.  After that nothing is caught: an error inside the exchange is a real error on every rank.
            if d not in self._symm:
                err = None
                try:
                    self._setup_fused(x_local, group, mode)
                except (RuntimeError, ImportError, AttributeError) as exc:
                    err = exc
                if not _all_ranks_ok(err is None, x_local.device, group):
                    self._symm.pop(d, None)
                    if mode != "auto":
                        raise RuntimeError(f"TCGNN_EXCHANGE={mode}: fused exchange unavailable on some rank ({err})")
                    self._symm_unavailable = True
                    print(f"[tcgnn sharding] rank {self.rank}: fused exchange unavailable ({err}); every rank uses "
                          f"the NCCL all-gather", file=sys.stderr, flush=True)
            if not self._symm_unavailable:
                return self._fused_exchange(x_local.contiguous(), group, mode)
        if round_tf32 and self.num_rows > 0:
            import TCGNN
            x_local = TCGNN.round_tf32(x_local.contiguous())
        if out is None:
            if self._x_all is None or self._x_all.shape[1] != d or self._x_all.device != x_local.device \
                    or self._x_all.dtype != x_local.dtype:
                self._x_all = torch.empty(self.num_cols, d, dtype=x_local.dtype, device=x_local.device)
            out = self._x_all
        if self.world_size == 1:
            out.copy_(x_local)
            return out
        views = [out[self.bounds[g]:self.bounds[g + 1]] for g in range(self.world_size)]
        x_local = x_local.contiguous()
        if dist.get_backend(group) == "nccl":
            dist.all_gather(views, x_local, group=group)       # uneven sizes: grouped ncclBroadcast over NVLink
        else:
            views[self.rank].copy_(x_local)
            for g in range(self.world_size):                   # gloo has no uneven all-gather
                if self.panel_rows(g) > 0:
                    dist.broadcast(views[g], src=dist.get_global_rank(group, g) if group is not None else g,
                                   group=group)
        return out

    def _setup_fused(self, x_local: torch.Tensor, group, mode: str) -> None:
        import torch.distributed._symmetric_memory as symm_mem
        d = x_local.shape[1]
        grp = group if group is not None else dist.group.WORLD
        buf = symm_mem.empty((self.num_cols, d), dtype=torch.float32, device=x_local.device)
        hdl = symm_mem.rendezvous(buf, grp)
        peers = [hdl.get_buffer(p, (self.num_cols, d), torch.float32) for p in range(self.world_size)]
        use_mc = mode in ("auto", "multicast") and bool(getattr(hdl, "has_multicast_support", False)) \
            and int(hdl.multicast_ptr) != 0
        if mode == "multicast" and not use_mc:
            raise RuntimeError("TCGNN_EXCHANGE=multicast but the group has no NVSwitch multicast support")
        self._symm[d] = (buf, hdl, peers, use_mc)

    def _fused_exchange(self, x_local: torch.Tensor, group, mode: str) -> torch.Tensor:
        import TCGNN
        d = x_local.shape[1]
        buf, hdl, peers, use_mc = self._symm[d]
        row_bytes = d * 4
        w, me = self.world_size, self.rank
        hdl.barrier(channel=0)                 # every rank has finished reading the previous gathered matrix
        if use_mc:
            if self.num_rows > 0:
                TCGNN.round_tf32_into(x_local, int(hdl.multicast_ptr) + self.row_base * row_bytes, d, True)
        elif mode == "p2p" or (w <= 2 and mode != "p2p2"):
            # direct: the rounding kernel writes the panel into every copy, one pass per peer (rotated)
            if self.num_rows > 0:
                for k in range(w):
                    p = (me + k) % w
                    TCGNN.round_tf32_into(x_local, peers[p].data_ptr() + self.row_base * row_bytes, d, False)
        else:
            # two-phase ("p2p2", default without multicast): panels are balanced on TC blocks, so their row counts
            # -- the bytes a rank has to send to EVERY peer -- differ 2-3x and the largest panel's owner bounds
            # the direct exchange.  Phase A: round the panel and scatter chunk j of it into rank j's copy (1/N of
            # the panel per peer).  Phase B: every rank now holds chunk `me` of every panel (~1/N of the matrix,
            # whatever the panel sizes) and pushes it to all peers in one launch: equal egress on every GPU.
            if self.num_rows > 0:
                for k in range(w):
                    j = (me + k) % w
                    c0, c1 = self._chunk(me, j)
                    if c1 > c0:
                        TCGNN.round_tf32_into(x_local[c0 - self.row_base:c1 - self.row_base],
                                              peers[j].data_ptr() + c0 * row_bytes, d, False)
            hdl.barrier(channel=2)
            begins, ends = [], []
            for g in range(w):                 # chunk `me` of every panel, my own included
                c0, c1 = self._chunk(g, me)
                if c1 > c0:
                    begins.append(c0)
                    ends.append(c1)
            others = [peers[(me + k) % w].data_ptr() for k in range(1, w)]
            if begins:
                TCGNN.push_rows(buf, others, begins, ends)
        hdl.barrier(channel=1)                 # every rank's rows have landed in every copy
        return buf

    def _chunk(self, g: int, j: int):
        """Global row range of chunk j (of world_size) of panel g."""
        b0, rows = self.bounds[g], self.bounds[g + 1] - self.bounds[g]
        return b0 + rows * j // self.world_size, b0 + rows * (j + 1) // self.world_size

    # ------------------------------------------------------------------ compute (GPU only)
    def spmm(self, x_all: torch.Tensor, edge_attention: Optional[torch.Tensor] = None,
             x_is_tf32: bool = False) -> torch.Tensor:
        import TCGNN
        if self.num_rows == 0:
            return x_all.new_zeros((0, x_all.shape[1]))
        if edge_attention is None:
            return TCGNN.panel_forward(x_all, self.row_base, *self.graph, x_is_tf32=x_is_tf32)[0]
        rp, ci, bp, e2c, e2r = self.graph
        return TCGNN.panel_forward_AGNN(x_all, self.row_base, rp, ci, edge_attention, bp, e2c, e2r,
                                        x_is_tf32=x_is_tf32)[0]

    def sddmm(self, x_all: torch.Tensor, x_is_tf32: bool = False) -> torch.Tensor:
        import TCGNN
        if self.num_rows == 0:
            return x_all.new_zeros((0,))
        return TCGNN.panel_forward_ef(x_all, self.row_base, *self.graph, x_is_tf32=x_is_tf32)[0]

    def _can_preround(self, x_local: torch.Tensor) -> bool:
        return x_local.is_cuda and x_local.shape[1] % 4 == 0

    def aggregate(self, x_local: torch.Tensor, group=None) -> torch.Tensor:
        """GCN/GIN/SAG aggregation of one layer -> the panel of A.X.  On NCCL groups (TCGNN_EXCHANGE=auto|overlap)
        the exchange is overlapped with the kernels source panel by source panel (`aggregate_overlapped`); else
        all-gather + one panel SpMM."""
        mode = os.environ.get("TCGNN_EXCHANGE", "auto")
        pre = self._can_preround(x_local)
        # 2 GPUs: the exchange is 0.15 ms of a 1.8 ms step and splitting the product in two costs more than hiding it
        # saves (1.83 vs 1.77 ms, profiles/r02h_exchange_probe2_n2.txt); from 3 GPUs on the overlapped path wins
        want_overlap = mode == "overlap" or (mode == "auto" and self.world_size > 2)
        if (pre and want_overlap and self.world_size > 1 and not self._ovl_unavailable
                and dist.get_backend(group) == "nccl"):
            d = x_local.shape[1]
            if d not in self._ovl:
                err = None
                try:
                    self._setup_overlap(d, x_local.device, group)
                except (RuntimeError, ImportError, AttributeError) as exc:
                    err = exc
                if not _all_ranks_ok(err is None, x_local.device, group):
                    self._ovl.pop(d, None)
                    if mode == "overlap":
                        raise RuntimeError(f"TCGNN_EXCHANGE=overlap unavailable on some rank ({err})")
                    self._ovl_unavailable = True
                    print(f"[tcgnn sharding] rank {self.rank}: overlapped exchange unavailable ({err}); every rank "
                          f"gathers first", file=sys.stderr, flush=True)
            if not self._ovl_unavailable:
                return self.aggregate_overlapped(x_local)
        return self.spmm(self.all_gather(x_local, group, round_tf32=pre), x_is_tf32=pre)

    # ------------------------------------------------------------------ overlapped exchange (GCN path)
    def build_source_subgraphs(self, dense_fraction: Optional[float] = None):
        """Split the panel's graph by SOURCE panel p: the edges whose column lies in panel p, as a CSR over the
        panel's rows whose column ids index the rows that will be shipped -- all of panel p (`dense`: ids rebased to
        the panel) or only the sorted unique referenced rows `ref` (ids = rank in `ref`).  Pure tensor code (runs on
        the CPU in the gloo tests); the SGT of every sub-graph comes from TCGNN.preprocess_panel."""
        if self._sub is not None:
            return self._sub
        import TCGNN
        if dense_fraction is None:   # ship a source's whole panel when at least this share of its rows is referenced
            dense_fraction = float(os.environ.get("TCGNN_DENSE_FRACTION", "0.7"))
        ci = self.column_index.long()
        e2r = self.edgeToRow.long()
        dev = ci.device
        subs = []
        for p in range(self.world_size):
            b0, b1 = self.bounds[p], self.bounds[p + 1]
            mask = (ci >= b0) & (ci < b1)
            cols = ci[mask]
            counts = torch.bincount(e2r[mask], minlength=self.num_rows)
            rp = torch.zeros(self.num_rows + 1, dtype=torch.int64, device=dev)
            torch.cumsum(counts, 0, out=rp[1:])
            ref = torch.unique(cols)                                   # sorted global ids
            dense = p == self.rank or ref.numel() >= dense_fraction * max(b1 - b0, 1)
            if dense:
                local = cols - b0
                n_src = b1 - b0
                ref_rows = None
            else:
                local = torch.searchsorted(ref, cols)
                n_src = int(ref.numel())
                ref_rows = (ref - b0).to(torch.int32)                  # rows of panel p, relative to its first row
            rp32 = rp.to(torch.int32).contiguous()
            ci32 = local.to(torch.int32).contiguous()
            nwin = (self.num_rows + BLK_H - 1) // BLK_H
            bp = torch.zeros(nwin, dtype=torch.int32, device=dev)
            e2c = torch.zeros(max(ci32.numel(), 1), dtype=torch.int32, device=dev)[:ci32.numel()]
            e2r_s = torch.zeros(max(ci32.numel(), 1), dtype=torch.int32, device=dev)[:ci32.numel()]
            if self.num_rows > 0:
                TCGNN.preprocess_panel(ci32, rp32, self.num_rows, max(n_src, 1), BLK_H, BLK_W, bp, e2c, e2r_s)
            subs.append({"graph": (rp32, ci32, bp, e2c, e2r_s), "n_src": n_src, "ref_rows": ref_rows,
                         "dense": dense, "edges": int(ci32.numel()), "rows": e2r[mask], "local": local})
        self._sub = subs
        return subs

    def arrival_order(self, q: Optional[int] = None) -> List[int]:
        """Sources in the order their pushes reach rank q: every rank pushes to its nearest successor first."""
        q = self.rank if q is None else q
        return [(q - k) % self.world_size for k in range(1, self.world_size)]

    def default_groups(self) -> List[int]:
        """How many consecutive arrivals share one product.  One product per source is the finest overlap, but a short
        launch pays its start-up, its tail and the sparser-window pipeline shape every time: at 8 GPUs eight products
        took 0.67 ms for 0.36 ms of kernel work (profiles/r02h_*parallel_products*).  Few groups, later ones larger
        (they wait longer anyway)."""
        env = os.environ.get("TCGNN_EXCHANGE_GROUPS")
        n = self.world_size - 1
        if env:
            sizes = [int(v) for v in env.split(",") if v.strip()]
            if sum(sizes) == n and all(v > 0 for v in sizes):
                return sizes
        if n <= 0:
            return []
        # measured on B200 (profiles/r02h_*): up to 4 GPUs one product over all remote sources wins (0.99 vs 1.05-1.08
        # ms at 4), at 8 GPUs one product per source (0.67 vs 0.72-0.94 ms)
        return [n] if self.world_size <= 4 else [1] * n

    def build_group_subgraphs(self, sizes: Optional[Sequence[int]] = None):
        """Merge the per-source sub-graphs of consecutive arrivals into one sub-graph per group: its column space is
        the concatenation of the rows those sources ship, in arrival order -- exactly how they sit in the receive
        area.  Returns a list of {"sources", "graph", "ncols", "edges"}."""
        import TCGNN
        sizes = list(sizes) if sizes is not None else self.default_groups()
        if self._groups is not None and self._groups[0] == tuple(sizes):
            return self._groups[1]
        if self._sub is not None and self._sub and "rows" not in self._sub[0]:
            self._sub = None                  # the raw edge lists were dropped after an earlier merge: rebuild
        subs = self.build_source_subgraphs()
        order = self.arrival_order()
        groups, i = [], 0
        nwin = (self.num_rows + BLK_H - 1) // BLK_H
        for sz in sizes:
            srcs = order[i:i + sz]
            i += sz
            rows, cols, off = [], [], 0
            for p_ in srcs:
                sb = subs[p_]
                rows.append(sb["rows"])
                cols.append(sb["local"] + off)
                off += sb["n_src"]
            ncols = max(off, 1)
            if len(srcs) == 1:
                g = subs[srcs[0]]["graph"]
            else:
                r = torch.cat(rows) if rows else torch.zeros(0, dtype=torch.int64)
                c = torch.cat(cols) if cols else torch.zeros(0, dtype=torch.int64)
                key = torch.sort(r * ncols + c).values
                r = key // ncols
                dev = key.device
                rp = torch.zeros(self.num_rows + 1, dtype=torch.int64, device=dev)
                torch.cumsum(torch.bincount(r, minlength=self.num_rows), 0, out=rp[1:])
                rp32 = rp.to(torch.int32).contiguous()
                ci32 = (key - r * ncols).to(torch.int32).contiguous()
                bp = torch.zeros(nwin, dtype=torch.int32, device=dev)
                e2c = torch.zeros(max(ci32.numel(), 1), dtype=torch.int32, device=dev)[:ci32.numel()]
                e2r_s = torch.zeros(max(ci32.numel(), 1), dtype=torch.int32, device=dev)[:ci32.numel()]
                if self.num_rows > 0:
                    TCGNN.preprocess_panel(ci32, rp32, self.num_rows, ncols, BLK_H, BLK_W, bp, e2c, e2r_s)
                g = (rp32, ci32, bp, e2c, e2r_s)
            groups.append({"sources": srcs, "graph": g, "ncols": off, "edges": int(g[1].numel())})
        self._groups = (tuple(sizes), groups)
        return groups

    def _setup_overlap(self, d: int, device, group) -> None:
        import torch.distributed._symmetric_memory as symm_mem
        w, me = self.world_size, self.rank
        grp = group if group is not None else dist.group.WORLD
        subs = self.build_source_subgraphs()
        # need[q][p] = rows rank q wants from rank p (0 on the diagonal); every rank learns the whole matrix
        mine = torch.tensor([0 if p == me else subs[p]["n_src"] for p in range(w)], dtype=torch.int64, device=device)
        need = [torch.zeros(w, dtype=torch.int64, device=device) for _ in range(w)]
        dist.all_gather(need, mine, group=group)
        need = torch.stack(need).cpu()
        # the row lists go to their sources (empty = "your whole panel")
        send = [torch.zeros(0, dtype=torch.int32, device=device) if (p == me or subs[p]["dense"])
                else subs[p]["ref_rows"].to(device) for p in range(w)]
        cnt_out = torch.tensor([t.numel() for t in send], dtype=torch.int64, device=device)
        cnt_in = torch.zeros(w, dtype=torch.int64, device=device)
        dist.all_to_all_single(cnt_in, cnt_out, group=group)
        cnt_in_l = [int(v) for v in cnt_in.cpu()]
        recv = torch.zeros(max(sum(cnt_in_l), 1), dtype=torch.int32, device=device)[:sum(cnt_in_l)]
        dist.all_to_all_single(recv, torch.cat(send) if w else recv, output_split_sizes=cnt_in_l,
                               input_split_sizes=[int(v) for v in cnt_out.cpu()], group=group)
        lists = list(torch.split(recv, cnt_in_l))           # lists[q] = my rows rank q wants (empty: all of them)
        # receive area: the sources' rows back to back in rank order; two copies (step parity) so a fast peer's
        # next push never lands in the rows a slow rank is still reading
        # receiver q lays its sources out in ARRIVAL order, so consecutive arrivals form one contiguous block
        offs = torch.zeros(w, w, dtype=torch.int64)
        totals = []
        for q in range(w):
            o = 0
            for p_ in self.arrival_order(q):
                offs[q, p_] = o
                o += int(need[q, p_])
            totals.append(o)
        rows_max = max(totals)
        st = OverlapState()
        st.groups = self.build_group_subgraphs()
        for sb in subs:                       # the raw edge lists were only needed for the merge
            sb.pop("rows", None)
            sb.pop("local", None)
        st.need, st.offs = need, offs
        st.recv = symm_mem.empty((2, max(rows_max, 1), d), dtype=torch.float32, device=device)
        st.recv_hdl = symm_mem.rendezvous(st.recv, grp)
        st.flags = symm_mem.empty((w,), dtype=torch.int32, device=device)
        st.flags.zero_()
        st.flags_hdl = symm_mem.rendezvous(st.flags, grp)
        st.peer_recv = [st.recv_hdl.get_buffer(q, (2, max(rows_max, 1), d), torch.float32) for q in range(w)]
        st.peer_flags = [st.flags_hdl.get_buffer(q, (w,), torch.int32) for q in range(w)]
        st.lists = lists
        st.stage = [torch.empty((int(need[q, me]), d), dtype=torch.float32, device=device)
                    if (q != me and lists[q].numel() > 0) else None for q in range(w)]
        st.xr = [torch.empty((self.num_rows, d), dtype=torch.float32, device=device) for _ in range(2)]
        st.step_dev = torch.zeros(1, dtype=torch.int32, device=device)   # step number; flag writes copy it
        st.err = torch.zeros(1, dtype=torch.int32, device=device)
        st.copy_stream = torch.cuda.Stream(device=device)
        st.product_streams = [torch.cuda.Stream(device=device) for _ in range(max(1, min(w - 1, 7)))]
        st.steps = 0
        st.graphs = {}
        torch.cuda.synchronize(device)
        st.flags_hdl.barrier(channel=0)                    # flags are zeroed everywhere before anybody pushes
        self._ovl[d] = st

    def aggregate_overlapped(self, x_local: torch.Tensor) -> torch.Tensor:
        """One layer's aggregation with the exchange hidden behind the kernels.

        This rank rounds its panel once, then its copy engines push the panel (or the packed rows a destination
        references) into every peer's receive area, nearest rank first, each push followed by a 4-byte flag write;
        meanwhile its SMs compute the own-panel product, then -- in the order the pushes of the other ranks arrive
        -- one TCGNN_ACCUMULATE product per source panel, each behind a stream-ordered wait on that source's flag.
        No barrier: the receive area is double-buffered by step parity, and a rank's pushes of step k + 1 are
        ordered after its own kernels of step k, which needed everybody's data of step k.

        The step is a fixed sequence of ~30 small launches and copies (at 8 GPUs the kernels of a reddit-sized step
        take 0.4 ms, issuing them from Python 0.9 ms), so after two eager steps it is captured once per step parity
        and input buffer as a CUDA graph and replayed (TCGNN_EXCHANGE_GRAPH=0 keeps it eager).  Nothing in it
        depends on host-side state: the step number lives in device memory (bumped inside the step), the flag
        writes copy it, the waits compare against it."""
        d = x_local.shape[1]
        st = self._ovl[d]
        if self.num_rows == 0 and self.world_size == 1:
            return x_local.new_zeros((0, d))
        st.steps += 1
        b = st.steps & 1
        x_local = x_local.contiguous()
        if os.environ.get("TCGNN_EXCHANGE_GRAPH", "1") == "0" or st.steps <= 2:
            return self._overlap_step(x_local, st, b)
        key = (x_local.data_ptr(), tuple(x_local.shape), b)
        g = st.graphs.get(key)
        if g is None:
            if len(st.graphs) >= 16:          # inputs that keep moving: stay eager instead of capturing forever
                return self._overlap_step(x_local, st, b)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                out = self._overlap_step(x_local, st, b)
            g = st.graphs[key] = (graph, out, x_local)   # x_local kept alive: its address is baked into the graph
        g[0].replay()
        return g[1].clone()

    def _overlap_step(self, x_local: torch.Tensor, st, b: int) -> torch.Tensor:
        """The step itself (eager or under stream capture): forks the copy stream off the current one, joins it."""
        import TCGNN
        d = x_local.shape[1]
        subs = self._sub
        w, me = self.world_size, self.rank
        cur = torch.cuda.current_stream(x_local.device)
        xr = st.xr[b]
        st.step_dev.add_(1)                                  # this step's number, on the device
        if self.num_rows > 0:
            TCGNN.round_tf32_into(x_local, xr.data_ptr(), d, False)
        cs = st.copy_stream
        cs.wait_stream(cur)
        packed = [q for q in ((me + k) % w for k in range(1, w)) if st.lists[q].numel() > 0 and int(st.need[q, me]) > 0]
        with torch.cuda.stream(cs):
            # pack first, all destinations: the pack kernels need SMs, which the products below keep busy -- packs
            # that queue behind them would delay every later push
            for q in packed:
                TCGNN.gather_rows(xr, st.lists[q], st.stage[q])
        if packed:
            cur.wait_stream(cs)
        with torch.cuda.stream(cs):
            for k in range(1, w):
                q = (me + k) % w
                n = int(st.need[q, me])
                if n > 0:
                    src = st.stage[q] if st.lists[q].numel() > 0 else xr
                    o = int(st.offs[q, me])
                    st.peer_recv[q][b, o:o + n].copy_(src, non_blocking=True)
                st.peer_flags[q][me:me + 1].copy_(st.step_dev, non_blocking=True)
        timeout_ms = int(os.environ.get("TCGNN_FLAG_TIMEOUT_MS", "20000"))
        # Every product ADDS into a zeroed Y (fp32 reduce-adds commute), so the products are independent of each
        # other: each group of sources' waits + product sit on their own stream, and the block scheduler fills an SM with the
        # next ready product's CTA the moment the previous product's CTA on it retires.  Issued back to back on one
        # stream the eight short launches of an 8-GPU step each paid their own start-up and tail (0.74 ms for 0.36 ms
        # of kernel work, profiles/r02h_*serial_products*).
        y = torch.zeros((self.num_rows, d), dtype=torch.float32, device=x_local.device)
        if self.num_rows > 0:
            TCGNN.source_forward(xr, *subs[me]["graph"], x_is_tf32=True, accumulate_into=y)
        branches = st.product_streams
        for s_ in branches:
            s_.wait_stream(cur)
        for gi, grp_ in enumerate(st.groups):
            with torch.cuda.stream(branches[gi % len(branches)]):
                for p in grp_["sources"]:
                    TCGNN.stream_wait_flag_dev(st.flags, p, st.step_dev, timeout_ms, st.err)
                n = grp_["ncols"]
                if self.num_rows > 0 and n > 0 and grp_["edges"] > 0:
                    o = int(st.offs[me, grp_["sources"][0]])
                    TCGNN.source_forward(st.recv[b, o:o + n], *grp_["graph"], x_is_tf32=True, accumulate_into=y)
        for s_ in branches:
            cur.wait_stream(s_)
        cur.wait_stream(cs)                                   # join: the next step bumps the counter the flag copies read
        return y

    def overlap_check(self) -> None:
        """Raises if a flag wait of the overlapped exchange timed out (synchronises; call outside timed regions)."""
        for st in self._ovl.values():
            if int(st.err.item()) != 0:
                raise RuntimeError("overlapped exchange: a source panel's rows did not arrive within "
                                   "TCGNN_FLAG_TIMEOUT_MS; the result of that step is incomplete")

    def overlap_stats(self, d: int):
        """Rows / bytes this rank receives per step in the overlapped exchange vs a full all-gather."""
        st = self._ovl.get(d)
        if st is None:
            return None
        rows = int(st.need[self.rank].sum())
        full = self.num_cols - self.num_rows
        return {"recv_rows": rows, "full_gather_rows": full, "recv_bytes": rows * d * 4,
                "dense_sources": sum(1 for p, sb in enumerate(self._sub) if p != self.rank and sb["dense"]),
                "packed_sources": sum(1 for p, sb in enumerate(self._sub) if p != self.rank and not sb["dense"])}

    def agnn_aggregate(self, x_local: torch.Tensor, attention_w: torch.Tensor, group=None):
        """AGNN aggregation (reference gnn_conv.py:125-132) on the panel: one gather serves SDDMM and
        the weighted SpMM.  Returns (Y_panel, edge_feature_panel)."""
        pre = self._can_preround(x_local)
        x_all = self.all_gather(x_local, group, round_tf32=pre)
        ef = self.sddmm(x_all, x_is_tf32=pre)
        att = torch.mm(ef.unsqueeze(-1), attention_w).transpose(0, 1).contiguous()
        return self.spmm(x_all, att, x_is_tf32=pre), ef
