#!/usr/bin/env bash
mkdir -p gpurun_out
for f in test_gpu_spmm test_gpu_sddmm test_gpu_vs_reference test_gpu_layers; do
  echo "=== $f"
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 300 -x 2>&1 | tail -40 > gpurun_out/$f.log
  tail -3 gpurun_out/$f.log
done
for wl in reddit-like-uniform reddit-like-rmat; do
for ab in 0 1 5 13; do
  TCGNN_ABLATE=$ab timeout 300 python tools/quick.py --workload $wl --iters 3 --tag ablate$ab 2>&1 | tail -1
done; done | tee gpurun_out/ablate4.txt
for op in sddmm wspmm; do timeout 300 python tools/quick.py --workload reddit-like-uniform --op $op --iters 3 2>&1 | tail -1; done | tee -a gpurun_out/ablate4.txt
timeout 300 python tools/quick.py --workload products-like-rmat --iters 3 2>&1 | tail -1 | tee -a gpurun_out/ablate4.txt
timeout 300 python tools/quick.py --workload products-like-rmat --op sddmm --iters 3 2>&1 | tail -1 | tee -a gpurun_out/ablate4.txt
timeout 300 python tools/quick.py --workload citeseer-like --iters 20 2>&1 | tail -1 | tee -a gpurun_out/ablate4.txt
