#!/usr/bin/env bash
mkdir -p gpurun_out
for wl in reddit-like-rmat reddit-like-uniform products-like-rmat rmat-10m-200m; do
for wc in 3 12 32 64 128; do
  TCGNN_WIN_COST=$wc timeout 300 python tools/quick.py --workload $wl --op spmm --iters 3 --tag wincost$wc 2>&1 | tail -1
done; done | tee gpurun_out/wincost_r02.txt
for wl in reddit-like-rmat products-like-rmat; do
for wc in 3 32 128; do
  TCGNN_WIN_COST=$wc timeout 300 python tools/quick.py --workload $wl --op wspmm_tile --iters 3 --tag wincost$wc 2>&1 | tail -1
done; done | tee -a gpurun_out/wincost_r02.txt
