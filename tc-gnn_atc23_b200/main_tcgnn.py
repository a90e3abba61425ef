#!/usr/bin/env python3
"""Training / single-kernel driver with the reference's command line and log lines
(reference main_tcgnn.py:18-181): `--dataset --dim --num_layers --hidden --classes --epochs --model
{gcn,gin,agnn} --single_kernel`; prints `Prep. (ms):`, `Train (ms):` and `=> SAG profiling avg (ms):`
in the reference's formats so its log scrapers (1_log2csv.py) keep working.

Additions: `--dataset` also accepts the synthetic specs of dataset.py (the reference's graphs are not
available offline), `--prep {cpu,gpu}` selects the host (multi-threaded) or device SGT, `--seed`.
"""
from __future__ import annotations

import argparse
import os.path as osp
import time

import torch
import torch.nn as nn
import torch.nn.functional as F

import TCGNN
from config import BLK_H, BLK_W
from dataset import TCGNN_dataset
from gnn_conv import SAG, AGNNConv, GCNConv, GINConv

CONVS = {"gcn": GCNConv, "gin": GINConv, "agnn": AGNNConv}


def parse_args(argv=None):
    p = argparse.ArgumentParser()
    p.add_argument("--dataset", type=str, default="amazon0601", help="dataset")
    p.add_argument("--dim", type=int, default=96, help="input embedding dimension")
    p.add_argument("--num_layers", type=int, default=2, help="num layers")
    p.add_argument("--hidden", type=int, default=16, help="hidden dimension")
    p.add_argument("--classes", type=int, default=22, help="number of output classes")
    p.add_argument("--epochs", type=int, default=200, help="number of epoches")
    p.add_argument("--model", type=str, default="gcn", help="GNN model", choices=sorted(CONVS))
    p.add_argument("--single_kernel", action="store_true", help="whether to profile a single SAG kernel")
    p.add_argument("--prep", type=str, default="cpu", choices=["cpu", "gpu"], help="where the SGT runs")
    p.add_argument("--seed", type=int, default=None, help="seed for synthetic graphs / features")
    return p.parse_args(argv)


def preprocess(dataset, where="cpu"):
    """SGT: returns the five int32 graph tensors on the GPU and the time spent (reference
    main_tcgnn.py:44-60)."""
    num_nodes, num_edges = dataset.num_nodes, dataset.num_edges
    column_index, row_pointers = dataset.column_index, dataset.row_pointers
    if where == "gpu":
        column_index, row_pointers = column_index.cuda(), row_pointers.cuda()
    dev = column_index.device
    num_row_windows = (num_nodes + BLK_H - 1) // BLK_H
    edgeToColumn = torch.zeros(num_edges, dtype=torch.int, device=dev)
    edgeToRow = torch.zeros(num_edges, dtype=torch.int, device=dev)
    blockPartition = torch.zeros(num_row_windows, dtype=torch.int, device=dev)
    start = time.perf_counter()
    TCGNN.preprocess(column_index, row_pointers, num_nodes, BLK_H, BLK_W, blockPartition, edgeToColumn, edgeToRow)
    if where == "gpu":
        torch.cuda.synchronize()
    elapsed = time.perf_counter() - start
    graph = tuple(t.cuda() for t in (row_pointers, column_index, blockPartition, edgeToColumn, edgeToRow))
    return graph, elapsed


class Net(torch.nn.Module):
    """conv1 -> relu -> dropout -> [hidden convs + relu] -> conv2 -> log_softmax (reference
    main_tcgnn.py:75-140, identical for the three model kinds up to the conv class)."""

    def __init__(self, conv, in_dim, hidden, classes, num_layers, dataset, graph):
        super().__init__()
        self.conv1 = conv(in_dim, hidden)
        self.hidden_layers = nn.ModuleList([conv(hidden, hidden) for _ in range(num_layers - 2)])
        self.conv2 = conv(hidden, classes)
        self.relu = nn.ReLU()
        self._dataset = [dataset]   # not a sub-module
        self._graph = graph

    def forward(self):
        g = self._graph
        x = self.relu(self.conv1(self._dataset[0].x, *g))
        x = F.dropout(x, training=self.training)
        for layer in self.hidden_layers:
            x = self.relu(layer(x, *g))
        return F.log_softmax(self.conv2(x, *g), dim=1)


def main(argv=None):
    args = parse_args(argv)
    print(args)
    path = args.dataset if (":" in args.dataset or osp.exists(args.dataset)) else \
        osp.join("tcgnn-ae-graphs/", args.dataset + ".npz")
    dataset = TCGNN_dataset(path, args.dim, args.classes, load_from_txt=False, seed=args.seed)
    graph, prep = preprocess(dataset, args.prep)
    print("Prep. (ms):\t{:.3f}".format(prep * 1e3))

    if args.single_kernel:
        SAG(*graph).profile(dataset.x)
        return 0

    device = torch.device("cuda:0")
    dataset = dataset.to(device)
    model = Net(CONVS[args.model], dataset.num_features, args.hidden, dataset.num_classes, args.num_layers, dataset,
                graph).to(device)
    optimizer = torch.optim.Adam(model.parameters(), lr=0.01, capturable=True)

    def train():
        model.train()
        optimizer.zero_grad()
        loss = F.nll_loss(model(), dataset.y)
        loss.backward()
        optimizer.step()
        return loss

    for _ in range(1, 10):   # dry run
        train()
    torch.cuda.synchronize()
    start_train = time.perf_counter()
    for _ in range(1, args.epochs + 1):
        train()
    torch.cuda.synchronize()
    train_time = time.perf_counter() - start_train
    print("Train (ms):\t{:6.3f}".format(train_time * 1e3 / args.epochs))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
