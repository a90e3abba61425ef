"""Node reordering that raises the density of the SGT's 16x8 TC blocks.

The reference's dataset class carries a `reorder_flag` that nothing ever sets (dataset.py:24) and reports how many
TC blocks the SGT leaves per dataset (logs/16x8_reduction.csv); how many blocks a graph needs is decided by which 16
rows share a window: a window costs ceil(#distinct neighbours of its 16 rows / 8) blocks, so rows with overlapping
neighbour sets should sit together.  Every block costs the kernels 8 gathered feature rows, whatever its occupancy --
block count IS the gathered-byte count, the quantity the SpMM / SDDMM kernels are bound by.

Relabelling is symmetric (A' = P A P^T): the graph stays the same graph, features and labels are permuted once with
it (`TCGNN_dataset(..., reorder=...)`), and every layer then runs on the relabelled ids at no run-time cost.

Orders (all device-side tensor code, deterministic):
  "degree"   rows by descending degree: hub rows share hub neighbours;
  "minhash"  rows by the two smallest hashes of their neighbour ids (a 2-permutation MinHash signature): rows with a
             large Jaccard overlap get equal or adjacent signatures with high probability, ties by degree;
  "hub"      hub clustering: every row keyed by its highest-degree neighbour, so the rows that attach to the same
             hub -- and hence share at least that column, usually several -- share windows.
"""
from __future__ import annotations

import torch

from config import BLK_H, BLK_W

_M1, _M2 = 0x9E3779B1, 0x85EBCA77


def _rows_of_edges(row_ptr: torch.Tensor) -> torch.Tensor:
    n = row_ptr.numel() - 1
    deg = (row_ptr[1:] - row_ptr[:-1]).long()
    return torch.repeat_interleave(torch.arange(n, device=row_ptr.device), deg)


def count_tc_blocks(row_ptr: torch.Tensor, col_idx: torch.Tensor, blk_h: int = BLK_H, blk_w: int = BLK_W) -> int:
    """sum over windows of ceil(max(#distinct columns, 1) / blk_w) -- the SGT's TC_Blocks (TCGNN.cpp:216)."""
    n = row_ptr.numel() - 1
    nwin = (n + blk_h - 1) // blk_h
    if col_idx.numel() == 0:
        return nwin
    win = _rows_of_edges(row_ptr) // blk_h
    key = torch.unique(win * n + col_idx.long())
    distinct = torch.bincount(key // n, minlength=nwin)
    return int(((torch.clamp(distinct, min=1) + blk_w - 1) // blk_w).sum())


def naive_tc_blocks(row_ptr: torch.Tensor, col_idx: torch.Tensor, blk_h: int = BLK_H, blk_w: int = BLK_W) -> int:
    """Blocks of a plain 16x8 tiling of the adjacency matrix that hold at least one non-zero (the `origin` column of
    the reference's logs/16x8_reduction.csv, 3_cnt_TC_blk_SpMM.py)."""
    n = row_ptr.numel() - 1
    if col_idx.numel() == 0:
        return 0
    ncb = (n + blk_w - 1) // blk_w
    win = _rows_of_edges(row_ptr) // blk_h
    return int(torch.unique(win * ncb + col_idx.long() // blk_w).numel())


def node_order(row_ptr: torch.Tensor, col_idx: torch.Tensor, method: str = "minhash") -> torch.Tensor:
    """perm[new_id] = old_id."""
    n = row_ptr.numel() - 1
    dev = row_ptr.device
    deg = (row_ptr[1:] - row_ptr[:-1]).long()
    if method == "none" or n == 0:
        return torch.arange(n, device=dev)
    if method == "degree":
        return torch.argsort(-deg, stable=True)
    rows = _rows_of_edges(row_ptr)
    cols = col_idx.long()
    big = torch.iinfo(torch.int64).max
    if method == "minhash":
        h1 = (cols * _M1 + 12345) & 0x7FFFFFFF
        h2 = (cols * _M2 + 6789) & 0x7FFFFFFF
        m1 = torch.full((n,), big, dtype=torch.int64, device=dev).scatter_reduce(0, rows, h1, "amin")
        m2 = torch.full((n,), big, dtype=torch.int64, device=dev).scatter_reduce(0, rows, h2, "amin")
        key = torch.stack([m1, m2, -deg], dim=1)
    elif method == "hub":
        # highest-degree neighbour (ties: smallest id): max over edges of deg[col] * n + (n - 1 - col)
        score = deg[cols] * n + (n - 1 - cols)
        top = torch.full((n,), -1, dtype=torch.int64, device=dev).scatter_reduce(0, rows, score, "amax")
        hub = torch.where(top >= 0, n - 1 - top % n, torch.full_like(top, n))
        hub_deg = torch.where(top >= 0, top // n, torch.zeros_like(top))
        key = torch.stack([-hub_deg, hub, -deg], dim=1)
    else:
        raise ValueError(f"unknown reordering method {method!r}")
    # lexicographic sort of the key columns (stable sorts, last column first)
    order = torch.arange(n, device=dev)
    for c in range(key.shape[1] - 1, -1, -1):
        order = order[torch.argsort(key[order, c], stable=True)]
    return order


def relabel(row_ptr: torch.Tensor, col_idx: torch.Tensor, perm: torch.Tensor):
    """CSR of P A P^T for perm[new] = old: (row_ptr', col_idx') int32, columns sorted inside every row."""
    n = row_ptr.numel() - 1
    dev = row_ptr.device
    inv = torch.empty(n, dtype=torch.int64, device=dev)
    inv[perm] = torch.arange(n, device=dev)
    rows = inv[_rows_of_edges(row_ptr)]
    cols = inv[col_idx.long()]
    key = torch.sort(rows * n + cols).values
    new_rows = key // n
    rp = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    torch.cumsum(torch.bincount(new_rows, minlength=n), 0, out=rp[1:])
    return rp.to(torch.int32), (key - new_rows * n).to(torch.int32)


def reorder_graph(row_ptr: torch.Tensor, col_idx: torch.Tensor, method: str = "minhash"):
    """(row_ptr', col_idx', perm, report) -- report = TC blocks before / after like the reference's reduction logs."""
    before = count_tc_blocks(row_ptr, col_idx)
    perm = node_order(row_ptr, col_idx, method)
    rp, ci = relabel(row_ptr, col_idx, perm)
    after = count_tc_blocks(rp, ci)
    nnz = int(col_idx.numel())
    report = {"method": method, "tc_blocks_before": before, "tc_blocks_after": after,
              "nnz_per_block_before": nnz / max(before, 1), "nnz_per_block_after": nnz / max(after, 1),
              "reduction_pct": 100.0 * (before - after) / max(before, 1)}
    return rp, ci, perm, report
