"""Graph input for the TC-GNN operators: the reference's `TCGNN_dataset` interface
(reference dataset.py:9-122) -- same constructor, same attributes -- plus synthetic graphs.

`path` may be
  * a `.npz` file with `src_li`, `dst_li`, `num_nodes` (reference dataset.py:74-80),
  * a whitespace separated `src dst` text file (`load_from_txt=True`, reference dataset.py:47-66),
  * a synthetic spec `rmat:<nodes>:<nnz>[:seed]` / `uniform:<nodes>:<nnz>[:seed]`, or one of the
    names in `graphgen.WORKLOADS` (e.g. `reddit-like-rmat`) -- the reference's graphs are a download
    that is not available offline.  A path like `tcgnn-ae-graphs/<name>.npz` that does not exist
    falls back to the synthetic workload `<name>` / `<name>-like` when there is one.
CSR is built exactly as the reference does (scipy coo -> csr: duplicates merged, columns sorted,
int32), features are standard normal, labels all ones (reference dataset.py:115,122).

`reorder="hub" | "minhash" | "degree"` relabels the nodes once (reorder.py: symmetric relabelling that packs rows
with overlapping neighbour sets into the same 16-row windows, i.e. fewer TC blocks for the same graph) -- what the
reference's never-set `reorder_flag` (dataset.py:24) stands for.  `self.perm[new_id] = old_id`; features and labels
are generated for the relabelled ids (the reference's are random / constant, so nothing else has to move).
"""
from __future__ import annotations

import os
import time

import numpy as np
import torch

import graphgen
from config import func


def _device():
    return torch.device("cuda" if torch.cuda.is_available() else "cpu")


class TCGNN_dataset(torch.nn.Module):
    def __init__(self, path, dim, num_class, load_from_txt=True, verbose=False, seed=None, reorder=None):
        super().__init__()
        self.nodes = set()
        self.load_from_txt = load_from_txt
        self.num_nodes = 0
        self.num_edges = 0
        self.num_features = dim
        self.num_classes = num_class
        self.edge_index = None
        self.reorder_flag = reorder not in (None, "none", False)
        self.reorder_method = reorder if self.reorder_flag else None
        self.reorder_report = None
        self.perm = None
        self.verbose_flag = verbose
        self.avg_degree = -1
        self.avg_edgeSpan = -1
        self.seed = seed
        self.init_edges(path)
        self.init_embedding(dim)
        self.init_labels(num_class)
        dev = _device()
        n = self.num_nodes
        idx = torch.arange(n, device=dev)
        self.train_mask = idx < int(n * 1.0)
        self.val_mask = idx < int(n * 0.3)
        self.test_mask = idx < int(n * 0.1)

    # ------------------------------------------------------------------ graph
    @staticmethod
    def _synthetic_spec(path):
        name = os.path.basename(str(path))
        if name.endswith(".npz"):
            name = name[:-4]
        parts = str(path).split(":")
        if parts[0] in ("rmat", "uniform") and len(parts) >= 3:
            return parts[0], int(parts[1]), int(parts[2]), int(parts[3]) if len(parts) > 3 else 0
        for cand in (name, name + "-like", name + "-like-rmat"):
            if cand in graphgen.WORKLOADS:
                n, nnz, _, kind = graphgen.WORKLOADS[cand]
                return kind, n, nnz, 0
        return None

    def init_edges(self, path):
        start = time.perf_counter()
        spec = None if os.path.exists(str(path)) else self._synthetic_spec(path)
        if spec is not None:
            kind, n, nnz, seed = spec
            if self.seed is not None:
                seed = self.seed
            rp, ci = graphgen.synthetic_graph(n, nnz, kind=kind, seed=seed, device=_device())
            self.num_nodes = n
            self.row_pointers = rp.cpu()
            self.column_index = ci.cpu()
            self.num_edges = int(self.column_index.numel())
            deg = (self.row_pointers[1:] - self.row_pointers[:-1]).to(torch.float32)
            self.avg_degree = self.num_edges / max(self.num_nodes, 1)
            rows = torch.repeat_interleave(torch.arange(n, dtype=torch.int64), deg.long())
            self.avg_edgeSpan = float((rows - self.column_index.long()).abs().float().mean()) if self.num_edges else 0.0
            if self.verbose_flag:
                print("# Synthetic {} graph (s): {:.3f}".format(kind, time.perf_counter() - start))
        else:
            if self.load_from_txt:
                pairs = np.loadtxt(path, dtype=np.int64, ndmin=2)
                src_li, dst_li = pairs[:, 0], pairs[:, 1]
                self.num_nodes = int(max(src_li.max(), dst_li.max())) + 1 if len(src_li) else 0
            else:
                if not str(path).endswith(".npz"):
                    raise ValueError("graph file must be a .npz file")
                obj = np.load(path)
                src_li, dst_li = obj["src_li"], obj["dst_li"]
                self.num_nodes = int(obj["num_nodes"])
            self.num_edges = len(src_li)
            self.edge_index = np.stack([src_li, dst_li])
            self.avg_degree = self.num_edges / max(self.num_nodes, 1)
            self.avg_edgeSpan = float(np.mean(np.abs(np.subtract(src_li, dst_li)))) if self.num_edges else 0.0
            if self.verbose_flag:
                print("# Loading (s): {:.3f}".format(time.perf_counter() - start))
            from scipy.sparse import coo_matrix
            t0 = time.perf_counter()
            csr = coo_matrix((np.ones(self.num_edges, dtype=np.int32), self.edge_index),
                             shape=(self.num_nodes, self.num_nodes)).tocsr()
            csr.sort_indices()
            if self.verbose_flag:
                print("# Build CSR (s): {:.3f}".format(time.perf_counter() - t0))
            # Unlike the reference (which keeps the raw pair count, dataset.py:79), num_edges is the number of
            # STORED non-zeros after scipy merged duplicate pairs: it is what callers size edgeToColumn /
            # edgeToRow with, and the operators read exactly len(column_index) entries (longer arrays are accepted).
            self.column_index = torch.from_numpy(csr.indices.astype(np.int32))
            self.row_pointers = torch.from_numpy(csr.indptr.astype(np.int32))
            self.num_edges = int(self.column_index.numel())
        if self.reorder_flag and self.num_edges > 0:
            import reorder as _reorder
            t0 = time.perf_counter()
            dev = _device()
            rp, ci, perm, rep = _reorder.reorder_graph(self.row_pointers.to(dev), self.column_index.to(dev),
                                                       self.reorder_method)
            self.row_pointers, self.column_index, self.perm = rp.cpu(), ci.cpu(), perm.cpu()
            self.reorder_report = rep
            if self.verbose_flag:
                print("# Reorder ({}) (s): {:.3f}  TC blocks {} -> {} ({:.1f} %)".format(
                    self.reorder_method, time.perf_counter() - t0, rep["tc_blocks_before"], rep["tc_blocks_after"],
                    rep["reduction_pct"]))
        if self.verbose_flag:
            print("# nodes: {}".format(self.num_nodes))
            print("# avg_degree: {:.2f}".format(self.avg_degree))
            print("# avg_edgeSpan: {}".format(int(self.avg_edgeSpan)))
        degrees = (self.row_pointers[1:] - self.row_pointers[:-1]).to(torch.float32)
        self.degrees = torch.sqrt(torch.clamp(degrees, min=float(func(0)))).to(_device())

    # ------------------------------------------------------------------ features / labels
    def init_embedding(self, dim):
        if self.seed is None:
            self.x = torch.randn(self.num_nodes, dim, device=_device())
        else:
            self.x = graphgen.features(self.num_nodes, dim, seed=self.seed, device=_device())

    def init_labels(self, num_class):
        self.y = torch.ones(self.num_nodes, dtype=torch.long, device=_device())

    def to(self, device):
        self.x = self.x.to(device)
        self.y = self.y.to(device)
        self.degrees = self.degrees.to(device)
        self.train_mask = self.train_mask.to(device)
        self.val_mask = self.val_mask.to(device)
        self.test_mask = self.test_mask.to(device)
        return self
