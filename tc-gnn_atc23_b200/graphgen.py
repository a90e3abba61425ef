"""Seeded synthetic graphs in the reference's input format (int32 CSR, columns sorted and unique
per row -- what `dataset.py:94-104` of the reference produces with scipy coo -> csr).

The reference's datasets (tcgnn-ae-graphs/*.npz) are a download that is not available offline, so
the benchmark configurations of BASELINE.json are reproduced by shape: same node count, same
(approximate) number of stored non-zeros, uniform or R-MAT degree structure.  Everything runs on
whatever device the caller names (the 114 M-edge reddit-sized graph is built on the GPU in about a
second; the unit tests use the CPU).  No oracle code is involved.
"""
from __future__ import annotations

import torch

RMAT_ABC = (0.57, 0.19, 0.19)   # d = 0.05 (Graph500 parameters)

# name -> (num_nodes, stored non-zeros, feature width, kind); sizes from SURVEY.md section 8
WORKLOADS = {
    "cora-like": (2708, 10858, 16, "uniform"),
    "citeseer-like": (3327, 9464, 16, "uniform"),
    "reddit-like-rmat": (232965, 114615892, 128, "rmat"),
    "reddit-like-uniform": (232965, 114615892, 128, "uniform"),
    "products-like-rmat": (2449029, 123718280, 256, "rmat"),
    "rmat-10m-200m": (10000000, 200000000, 256, "rmat"),
}


def csr_from_pairs(src: torch.Tensor, dst: torch.Tensor, num_nodes: int):
    """(row_ptr int32[N+1], col_idx int32[nnz]) with duplicates merged and columns sorted."""
    key = torch.unique(src.to(torch.int64) * num_nodes + dst.to(torch.int64))   # sorted
    return csr_from_keys(key, num_nodes)


def csr_from_keys(key: torch.Tensor, num_nodes: int):
    rows = torch.div(key, num_nodes, rounding_mode="floor")
    cols = (key - rows * num_nodes).to(torch.int32)
    counts = torch.bincount(rows, minlength=num_nodes)
    row_ptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=key.device)
    torch.cumsum(counts, 0, out=row_ptr[1:])
    if int(row_ptr[-1]) > 2**31 - 1:
        raise ValueError("graph has more than 2^31-1 stored edges; the operator contract is int32 CSR")
    return row_ptr.to(torch.int32), cols.contiguous()


def _generator(seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def _uniform_pairs(m, num_nodes, gen, device):
    src = torch.randint(0, num_nodes, (m,), generator=gen, device=device, dtype=torch.int64)
    dst = torch.randint(0, num_nodes, (m,), generator=gen, device=device, dtype=torch.int64)
    return src, dst


def _rmat_pairs(m, num_nodes, gen, device, abc=RMAT_ABC):
    a, b, c = abc
    scale = max(1, (int(num_nodes) - 1).bit_length())
    src = torch.zeros(m, dtype=torch.int64, device=device)
    dst = torch.zeros(m, dtype=torch.int64, device=device)
    for _ in range(scale):
        r = torch.rand(m, generator=gen, device=device)
        src = (src << 1) | (r >= a + b).to(torch.int64)
        dst = (dst << 1) | (((r >= a) & (r < a + b)) | (r >= a + b + c)).to(torch.int64)
    return src % num_nodes, dst % num_nodes


def synthetic_graph(num_nodes: int, target_nnz: int, kind: str = "rmat", seed: int = 0, device="cpu",
                    symmetric: bool = True, fill: bool = True):
    """Seeded graph with about `target_nnz` stored non-zeros (never more).

    Pairs are drawn (uniformly or R-MAT), mirrored when `symmetric` (the reference's backward pass
    assumes A == A^T, gnn_conv.py:76-85) and de-duplicated.  R-MAT loses many pairs to duplicates,
    so with `fill` further batches are drawn until at least 99 % of the target is stored.
    Returns (row_ptr, col_idx) as int32 tensors on `device`."""
    if kind not in ("rmat", "uniform"):
        raise ValueError(f"unknown graph kind {kind!r}")
    device = torch.device(device)
    gen = _generator(seed, device)
    draw = _rmat_pairs if kind == "rmat" else _uniform_pairs
    keys = torch.empty(0, dtype=torch.int64, device=device)
    per_pair = 2 if symmetric else 1
    for _ in range(64):
        missing = target_nnz - keys.numel()
        if missing <= 0 or (keys.numel() > 0 and (not fill or keys.numel() >= 0.99 * target_nnz)):
            break
        m = max(missing // per_pair, 1)
        src, dst = draw(m, num_nodes, gen, device)
        k = src * num_nodes + dst
        if symmetric:
            k = torch.cat([k, dst * num_nodes + src])
        keys = torch.unique(torch.cat([keys, k]))
        del src, dst, k
    if keys.numel() > target_nnz:
        keys = _trim(keys, num_nodes, target_nnz, symmetric)
    return csr_from_keys(keys, num_nodes)


def _trim(keys: torch.Tensor, num_nodes: int, target_nnz: int, symmetric: bool) -> torch.Tensor:
    """Drop the overshoot from the tail.  On a symmetric graph an off-diagonal pair leaves together with its mirror
    (cutting `keys[:target]` would keep (a, b) without (b, a)); the result may be one entry short of the target."""
    excess = keys.numel() - target_nnz
    if not symmetric:
        return keys[:target_nnz]
    rows = torch.div(keys, num_nodes, rounding_mode="floor")
    cols = keys - rows * num_nodes
    lower = torch.nonzero(rows > cols).flatten()          # one representative per mirrored pair, sorted by key
    n_pairs = min((excess + 1) // 2, lower.numel())
    drop = lower[lower.numel() - n_pairs:]
    keep = torch.ones(keys.numel(), dtype=torch.bool, device=keys.device)
    keep[drop] = False
    mirror = cols[drop] * num_nodes + rows[drop]
    keep[torch.searchsorted(keys, mirror)] = False
    keys = keys[keep]
    if keys.numel() > target_nnz:                         # only self-loops were left to drop
        diag = torch.nonzero(torch.div(keys, num_nodes, rounding_mode="floor") == keys % num_nodes).flatten()
        keep = torch.ones(keys.numel(), dtype=torch.bool, device=keys.device)
        keep[diag[diag.numel() - min(diag.numel(), keys.numel() - target_nnz):]] = False
        keys = keys[keep]
    return keys


def workload(name: str, device="cuda", seed: int = 0):
    """(row_ptr, col_idx, num_nodes, feature_dim) of a named BASELINE.json-shaped workload."""
    n, nnz, dim, kind = WORKLOADS[name]
    rp, ci = synthetic_graph(n, nnz, kind=kind, seed=seed, device=device)
    return rp, ci, n, dim


def features(num_nodes: int, dim: int, seed: int = 0, device="cpu") -> torch.Tensor:
    """Standard-normal node features like the reference (`dataset.py:115`), but seeded."""
    return torch.randn(num_nodes, dim, generator=_generator(seed + 1000003, torch.device(device)), device=device)
