#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python tools/umma_bench.py 2>&1 | tee gpurun_out/umma_bench.txt
for wl in reddit-like-uniform; do
for ab in 0 8 9 12 13; do
  TCGNN_ABLATE=$ab timeout 300 python tools/quick.py --workload $wl --iters 3 --tag ablate$ab 2>&1 | tail -1
done; done | tee gpurun_out/ablate2.txt
