#!/usr/bin/env bash
mkdir -p gpurun_out
for f in test_gpu_sddmm test_gpu_fused_ops; do
  echo "=== $f"; timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 600 2>&1 | grep -v Warn | tail -30 > gpurun_out/$f.log; tail -2 gpurun_out/$f.log
done
TCGNN_SDDMM_TEAM=2 timeout 600 python tools/stress_sddmm.py products-like-rmat 256 20 2>&1 | grep -v Warn | tail -3
{
for wl in reddit-like-rmat reddit-like-uniform products-like-rmat rmat-10m-200m; do
  timeout 300 python tools/quick.py --workload $wl --op sddmm --iters 3 --tag prefetch 2>&1 | tail -1
  timeout 300 python tools/quick.py --workload $wl --op agnn --iters 3 --tag prefetch 2>&1 | tail -1
  timeout 300 python tools/quick.py --workload $wl --op spmm --iters 3 --tag shape0 2>&1 | tail -1
  TCGNN_SPMM_SHAPE=1 timeout 300 python tools/quick.py --workload $wl --op spmm --iters 3 --tag shape1 2>&1 | tail -1
  TCGNN_SPMM_SHAPE=1 timeout 300 python tools/quick.py --workload $wl --op wspmm_tile --iters 3 --tag shape1 2>&1 | tail -1
done; } | tee gpurun_out/timings_f.txt
TCGNN_SPMM_SHAPE=1 timeout 900 python -m pytest tests/test_gpu_spmm.py -m gpu -q --timeout 600 2>&1 | tail -2
