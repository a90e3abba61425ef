#!/usr/bin/env bash
mkdir -p gpurun_out
for f in test_gpu_spmm test_gpu_sddmm test_gpu_vs_reference test_gpu_layers test_gpu_sharding test_gpu_fullsize; do
  echo "=== $f"
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 600 -x 2>&1 | grep -v Warning | tail -40 > gpurun_out/$f.log
  tail -3 gpurun_out/$f.log
done
for wl in reddit-like-rmat reddit-like-uniform products-like-rmat rmat-10m-200m; do
for wc in 4 0 8 16; do
  TCGNN_WIN_COST=$wc timeout 300 python tools/quick.py --workload $wl --iters 3 --tag wincost$wc 2>&1 | tail -1
done; done | tee gpurun_out/wincost.txt
for wl in reddit-like-rmat reddit-like-uniform; do
for ps in 1 2 3 4 5; do
  TCGNN_PRESET=$ps timeout 300 python tools/quick.py --workload $wl --iters 3 --tag preset$ps 2>&1 | tail -1
done; done | tee -a gpurun_out/wincost.txt
for op in sddmm wspmm; do timeout 300 python tools/quick.py --workload reddit-like-uniform --op $op --iters 3 2>&1 | tail -1; done | tee -a gpurun_out/wincost.txt
timeout 300 python tools/quick.py --workload products-like-rmat --op sddmm --iters 3 2>&1 | tail -1 | tee -a gpurun_out/wincost.txt
timeout 300 python tools/quick.py --workload citeseer-like --iters 20 2>&1 | tail -1 | tee -a gpurun_out/wincost.txt
TCGNN_TRACE_CTA=127 TCGNN_TRACE=gpurun_out/trace_cta127.bin timeout 300 python tools/quick.py --workload rmat-10m-200m --iters 1 --tag cta127 2>&1 | tail -1
python tools/trace.py gpurun_out/trace_cta127.bin 50 500 | tail -14 | tee -a gpurun_out/wincost.txt
