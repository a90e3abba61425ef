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

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist

from config import BLK_H, BLK_W


def partition_rows(row_ptr: torch.Tensor, world_size: int, blk_h: int = BLK_H) -> List[int]:
    """Row boundaries [b_0 = 0, ..., b_world = N], each a multiple of `blk_h` (except N), balancing
    the stored non-zeros per panel.  Deterministic, identical on every rank."""
    if world_size < 1:
        raise ValueError("world_size must be >= 1")
    n = int(row_ptr.numel()) - 1
    rp = row_ptr.to(torch.int64).cpu()
    nnz = int(rp[-1] - rp[0])
    bounds = [0]
    for k in range(1, world_size):
        target = int(rp[0]) + (nnz * k) // world_size
        r = int(torch.searchsorted(rp, torch.tensor(target, dtype=torch.int64), right=False))
        r = min(max(r, 0), n)
        r = (r + blk_h // 2) // blk_h * blk_h          # nearest window boundary
        r = min(max(r, bounds[-1]), n // blk_h * blk_h if n >= blk_h else 0)
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
        self.bounds = list(bounds) if bounds is not None else partition_rows(row_ptr, world_size)
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
        again: pass the result to spmm / sddmm with x_is_tf32=True."""
        if x_local.shape[0] != self.num_rows:
            raise ValueError(f"x_local has {x_local.shape[0]} rows, the panel has {self.num_rows}")
        d = x_local.shape[1]
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
