#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 300 python tools/umma_bench.py 2>&1 | head -8 | tee gpurun_out/umma_bench.txt
for f in test_gpu_spmm test_gpu_sddmm; do
  echo "=== $f"
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 300 -x 2>&1 | tail -40 > gpurun_out/$f.log
  tail -3 gpurun_out/$f.log
done
for wl in reddit-like-uniform reddit-like-rmat; do
for lag in 2 3 4 5; do
  TCGNN_LAG=$lag timeout 300 python tools/quick.py --workload $wl --iters 3 --tag lag$lag 2>&1 | tail -1
done; done | tee gpurun_out/lag.txt
for ab in 1 4 5 8 13; do
  TCGNN_ABLATE=$ab timeout 300 python tools/quick.py --workload reddit-like-uniform --iters 3 --tag ablate$ab 2>&1 | tail -1
done | tee -a gpurun_out/lag.txt
timeout 300 python tools/quick.py --workload reddit-like-uniform --op sddmm --iters 3 2>&1 | tail -1 | tee -a gpurun_out/lag.txt
timeout 300 python tools/quick.py --workload products-like-rmat --iters 3 2>&1 | tail -1 | tee -a gpurun_out/lag.txt
