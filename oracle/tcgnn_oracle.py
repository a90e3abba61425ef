"""CPU oracle for the TC-GNN aggregation path -- TEST INFRASTRUCTURE ONLY.

This file is a NumPy restatement of what the reference computes on this path.  It is
imported only by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline /
--impl reference leg, and only as the checker -- the product (tc-gnn_atc23_b200/) never
imports it and fails loudly when its CUDA library is missing.

Parity pinning (see tests/golden/make_golden.py and tests/test_oracle.py):
  * SGT is pinned bit-exactly against the reference's own compiled `TCGNN.preprocess`
    (oracle/_ref, built from /root/reference/TCGNN_conv unmodified) on the graphs stored in
    tests/golden/sgt_*.npz.
  * SpMM / SDDMM: the reference ships no tests or golden vectors; they are pinned against
    (a) the reference's hand-checkable fixtures (gnn_conv.py:13-23 `gen_test_tensor`,
    gnn_conv.py:61 `ones_like`) and (b) the reference's own CUDA kernels run on the B200 box
    (tests/test_gpu_vs_reference.py, uses oracle/_ref on the GPU).

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import numpy as np

BLK_H = 16  # TCGNN_conv/config.h:4, config.py:1
BLK_W = 8   # TCGNN_conv/config.h:5, config.py:2


def tf32_rna(x: np.ndarray) -> np.ndarray:
    """`cvt.rna.tf32.f32` as used by wmma::__float_to_tf32 (TCGNN_kernel.cu:436-444;
    semantics /usr/local/cuda/include/crt/mma.h:96-103): round to nearest, ties away from
    zero, keep 10 mantissa bits.  Inf/NaN pass through."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32)
    special = (u & np.uint32(0x7F800000)) == np.uint32(0x7F800000)
    r = (u + np.uint32(0x00001000)) & np.uint32(0xFFFFE000)
    r = np.where(special, u, r)
    return r.view(np.float32).reshape(x.shape)


def sgt(row_pointers: np.ndarray, column_index: np.ndarray, num_nodes: int,
        blk_h: int = BLK_H, blk_w: int = BLK_W):
    """Sparse-graph translation, TCGNN.cpp:172-226 (+ inplace_deduplication :157-170).

    Returns (blockPartition[W], edgeToColumn[E], edgeToRow[E], printed_tc_blocks).
    Quirks that are part of the contract:
      * an empty window still yields blockPartition == 1 (the std::map is seeded with
        array[0] of a zero-length buffer, TCGNN.cpp:160),
      * the loop `iter < num_nodes + 1` (TCGNN.cpp:200) runs one extra, empty window when
        num_nodes % blk_h == 0: it adds 1 to the printed TC_Blocks and writes one int past
        the end of blockPartition.  The out-of-bounds write is NOT reproduced; the printed
        total is.
    """
    rp = np.asarray(row_pointers, dtype=np.int64)
    ci = np.asarray(column_index, dtype=np.int64)
    n = int(num_nodes)
    e = ci.shape[0]
    nwin = (n + blk_h - 1) // blk_h
    edge_to_row = np.zeros(e, dtype=np.int32)
    deg = rp[1:n + 1] - rp[:n]
    # TCGNN.cpp:194-197
    edge_to_row[rp[0]:rp[n]] = np.repeat(np.arange(n, dtype=np.int32), deg)
    block_partition = np.zeros(nwin, dtype=np.int32)
    edge_to_col = np.zeros(e, dtype=np.int32)
    for w in range(nwin):
        s = rp[w * blk_h]
        t = rp[min(w * blk_h + blk_h, n)]
        nb = ci[s:t].astype(np.uint32)            # memcpy as unsigned, TCGNN.cpp:205-206
        uniq = np.unique(nb)                      # sort + dedup, TCGNN.cpp:209-213
        block_partition[w] = (max(len(uniq), 1) + blk_w - 1) // blk_w   # TCGNN.cpp:216
        edge_to_col[s:t] = np.searchsorted(uniq, nb).astype(np.int32)   # TCGNN.cpp:220-223
    printed = int(block_partition.sum()) + (1 if n % blk_h == 0 else 0)
    return block_partition, edge_to_col, edge_to_row, printed


def _row_ids(row_pointers, n):
    rp = np.asarray(row_pointers, dtype=np.int64)
    return np.repeat(np.arange(n, dtype=np.int64), rp[1:n + 1] - rp[:n])


def spmm(x: np.ndarray, row_pointers, column_index, edge_weight=None, *, tf32: bool = True,
         dtype=np.float32) -> np.ndarray:
    """Neighbour aggregation Y = A.X.

    Unweighted (TCGNN_kernel.cu:336-454): A is the 0/1 pattern (sparse_A[..] = 1 at :405; a
    duplicated (row, col) pair would still be a single 1, hence the per-row unique), X is
    rounded with cvt.rna.tf32 (:441-443), fp32 accumulate.
    Weighted (TCGNN_kernel.cu:459-578): A[row, col] = edgeAttention[e] (:529), also rounded.
    `tf32=False, dtype=np.float64` gives the "true math" variant used for the 1e-3 bound.
    The reference leaves columns >= 128 and the D % 16 tail at zero; the
    oracle computes them, parity against the reference kernel is asserted on its defined region.
    """
    x = np.asarray(x)
    n = len(row_pointers) - 1
    rp = np.asarray(row_pointers, dtype=np.int64)
    ci = np.asarray(column_index, dtype=np.int64)
    rows = _row_ids(rp, n)
    xv = tf32_rna(x) if tf32 else x.astype(np.float32)
    xv = xv.astype(dtype)
    y = np.zeros((n, x.shape[1]), dtype=dtype)
    if edge_weight is None:
        # pattern semantics: drop duplicated (row, col) pairs
        key = rows * (int(ci.max()) + 1 if len(ci) else 1) + ci
        _, first = np.unique(key, return_index=True)
        np.add.at(y, rows[first], xv[ci[first]])
    else:
        w = np.asarray(edge_weight, dtype=np.float32).reshape(-1)
        wv = (tf32_rna(w) if tf32 else w).astype(dtype)
        np.add.at(y, rows, wv[:, None] * xv[ci])
    return y


def spmm_abs(x, row_pointers, column_index, edge_weight=None) -> np.ndarray:
    """sum of |terms| per output element -- the scale for the norm-wise parity bound."""
    w = None if edge_weight is None else np.abs(np.asarray(edge_weight, dtype=np.float32))
    return spmm(np.abs(np.asarray(x, dtype=np.float32)), row_pointers, column_index, w,
                tf32=False, dtype=np.float64)


def sddmm(x: np.ndarray, row_pointers, column_index, *, tf32: bool = True,
          dtype=np.float32) -> np.ndarray:
    """Edge scores out[e] = <X[row(e)], X[col(e)]>, TCGNN_kernel.cu:584-728: both operands
    rounded with cvt.rna.tf32 (:701-709), fp32 accumulate in k-steps of 8 (:604,667), only
    edge positions written (:719-726)."""
    n = len(row_pointers) - 1
    rows = _row_ids(row_pointers, n)
    ci = np.asarray(column_index, dtype=np.int64)
    xv = (tf32_rna(x) if tf32 else np.asarray(x, dtype=np.float32)).astype(dtype)
    out = np.zeros(len(ci), dtype=dtype)
    step = 1 << 20
    for s in range(0, len(ci), step):
        out[s:s + step] = np.einsum("ed,ed->e", xv[rows[s:s + step]], xv[ci[s:s + step]])
    return out


def sddmm_abs(x, row_pointers, column_index) -> np.ndarray:
    n = len(row_pointers) - 1
    rows = _row_ids(row_pointers, n)
    ci = np.asarray(column_index, dtype=np.int64)
    xa = np.abs(np.asarray(x, dtype=np.float64))
    out = np.zeros(len(ci), dtype=np.float64)
    step = 1 << 20
    for s in range(0, len(ci), step):
        out[s:s + step] = np.einsum("ed,ed->e", xa[rows[s:s + step]], xa[ci[s:s + step]])
    return out


# ---- layer-level restatements (gnn_conv.py) used by the autograd-wrapper tests -------------

def gcn_layer_forward(x, w, rp, ci):
    """gnn_conv.py:54-73: X' = X @ W ; Y = SpMM(X')."""
    return spmm((np.asarray(x, np.float32) @ np.asarray(w, np.float32)), rp, ci)


def agnn_layer_forward(x, w, attention_w, rp, ci):
    """gnn_conv.py:117-136: X' = X@W ; ef = SDDMM(X') ; att = (ef[:,None] @ attention_w).T ;
    Y = SpMM_w(X', att[0])."""
    xp = np.asarray(x, np.float32) @ np.asarray(w, np.float32)
    ef = sddmm(xp, rp, ci)
    att = (ef[:, None] @ np.asarray(attention_w, np.float32)).T.copy()
    return spmm(xp, rp, ci, att[0]), ef, att


# ---- synthetic graphs (dataset.py:94-104 builds CSR through scipy coo -> csr) -----------------

def csr_from_edges(src, dst, num_nodes):
    """dataset.py:94-104: coo_matrix((ones, (src, dst))).tocsr() -- duplicates are summed
    (so the pattern is unique per row) and columns come out sorted; int32 arrays."""
    from scipy.sparse import coo_matrix
    val = np.ones(len(src), dtype=np.int8)
    csr = coo_matrix((val, (np.asarray(src), np.asarray(dst))), shape=(num_nodes, num_nodes)).tocsr()
    csr.sum_duplicates()
    csr.sort_indices()
    return csr.indptr.astype(np.int32), csr.indices.astype(np.int32)


def random_graph(num_nodes, num_edges, seed=0, symmetric=True):
    rng = np.random.default_rng(seed)
    m = num_edges // 2 if symmetric else num_edges
    src = rng.integers(0, num_nodes, size=m, dtype=np.int64)
    dst = rng.integers(0, num_nodes, size=m, dtype=np.int64)
    if symmetric:
        src, dst = np.concatenate([src, dst]), np.concatenate([dst, src])
    return csr_from_edges(src, dst, num_nodes)


def rmat_graph(num_nodes, num_edges, seed=0, a=0.57, b=0.19, c=0.19, symmetric=True):
    """R-MAT (a,b,c,d)=(0.57,0.19,0.19,0.05) ids folded mod num_nodes."""
    rng = np.random.default_rng(seed)
    scale = max(1, int(np.ceil(np.log2(max(num_nodes, 2)))))
    m = num_edges // 2 if symmetric else num_edges
    src = np.zeros(m, dtype=np.int64)
    dst = np.zeros(m, dtype=np.int64)
    for _ in range(scale):
        r = rng.random(m)
        src = (src << 1) | (r >= a + b)
        dst = (dst << 1) | (((r >= a) & (r < a + b)) | (r >= a + b + c))
    src %= num_nodes
    dst %= num_nodes
    if symmetric:
        src, dst = np.concatenate([src, dst]), np.concatenate([dst, src])
    return csr_from_edges(src, dst, num_nodes)
