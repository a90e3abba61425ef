"""1-D row-panel (destination-node) sharding of the aggregation path across the GPUs of one box.

New work -- the reference is single process / single GPU.  Design (DESIGN.md "multi-GPU"):

  * rank g owns the rows [bounds[g], bounds[g+1]) of the graph; boundaries fall on multiples of
    BLK_H = 16 rows, so every 16-row window belongs to exactly one rank and a panel's SGT arrays are
    exactly the panel's slice of the whole graph's SGT arrays (window-local ranks, per-window tile
    counts) -- sharded results are bit-identical to the single-GPU ones;
  * boundaries are chosen on the prefix sum of the CSR row pointer so the stored non-zeros (= the
    gather work) per rank are balanced, not the row counts (R-MAT rows are heavily skewed);
  * column ids stay global.  Per layer the only exchange is ONE all-gather of the layer input
    (every rank contributes its panel of X, receives the others') over NCCL / NVLink; the output of
    SpMM stays sharded -- it is the rank's panel of the next layer's input;
  * the kernels take the gathered matrix plus the panel plan (`tcgnn_plan_create_panel`), AGNN's
    SDDMM + weighted SpMM share one gather.

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
                   block_partition: Optional[torch.Tensor] = None, window_cost: int = 3) -> List[int]:
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
            self.bounds = partition_rows(row_ptr, world_size, block_partition=sgt[0] if sgt is not None else None)
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
        if (round_tf32 and out is None and self.world_size > 1 and x_local.is_cuda and d % 4 == 0
                and mode != "nccl" and not self._symm_unavailable and dist.get_backend(group) == "nccl"):
            try:
                return self._fused_exchange(x_local.contiguous(), group, mode)
            except (RuntimeError, ImportError, AttributeError) as exc:
                # symmetric memory needs P2P access + a working rendezvous on every rank; the failure is collective
                # (same platform everywhere), so every rank falls back to the NCCL all-gather together
                if mode != "auto":
                    raise
                self._symm_unavailable = True
                print(f"[tcgnn sharding] rank {self.rank}: fused exchange unavailable ({exc}); using NCCL all-gather",
                      file=sys.stderr, flush=True)
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

    def _fused_exchange(self, x_local: torch.Tensor, group, mode: str) -> torch.Tensor:
        import TCGNN
        import torch.distributed._symmetric_memory as symm_mem
        d = x_local.shape[1]
        st = self._symm.get(d)
        if st is None:
            grp = group if group is not None else dist.group.WORLD
            buf = symm_mem.empty((self.num_cols, d), dtype=torch.float32, device=x_local.device)
            hdl = symm_mem.rendezvous(buf, grp)
            peers = [hdl.get_buffer(p, (self.num_cols, d), torch.float32) for p in range(self.world_size)]
            use_mc = mode in ("auto", "multicast") and bool(getattr(hdl, "has_multicast_support", False)) \
                and int(hdl.multicast_ptr) != 0
            if mode == "multicast" and not use_mc:
                raise RuntimeError("TCGNN_EXCHANGE=multicast but the group has no NVSwitch multicast support")
            st = self._symm[d] = (buf, hdl, peers, use_mc)
        buf, hdl, peers, use_mc = st
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
        """GCN/GIN/SAG aggregation of one layer: all-gather + panel SpMM -> the panel of A.X."""
        pre = self._can_preround(x_local)
        return self.spmm(self.all_gather(x_local, group, round_tf32=pre), x_is_tf32=pre)

    def agnn_aggregate(self, x_local: torch.Tensor, attention_w: torch.Tensor, group=None):
        """AGNN aggregation (reference gnn_conv.py:125-132) on the panel: one gather serves SDDMM and
        the weighted SpMM.  Returns (Y_panel, edge_feature_panel)."""
        pre = self._can_preround(x_local)
        x_all = self.all_gather(x_local, group, round_tf32=pre)
        ef = self.sddmm(x_all, x_is_tf32=pre)
        att = torch.mm(ef.unsqueeze(-1), attention_w).transpose(0, 1).contiguous()
        return self.spmm(x_all, att, x_is_tf32=pre), ef
