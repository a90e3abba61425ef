#!/usr/bin/env bash
mkdir -p gpurun_out
for ps in 0 3; do
  echo "=== test_gpu_spmm preset $ps"
  TCGNN_PRESET=$ps timeout 900 python -m pytest tests/test_gpu_spmm.py -m gpu -q --timeout 300 -x 2>&1 | tail -3
done
echo "=== other tests"
timeout 900 python -m pytest tests/test_gpu_layers.py tests/test_gpu_sharding.py tests/test_gpu_vs_reference.py tests/test_gpu_fullsize.py -m gpu -q --timeout 600 -x 2>&1 | tail -2
for wl in reddit-like-rmat reddit-like-uniform products-like-rmat rmat-10m-200m citeseer-like; do
for ps in 0 1 2 3 4 5 6; do
  TCGNN_PRESET=$ps timeout 300 python tools/quick.py --workload $wl --iters 3 --tag preset$ps 2>&1 | tail -1
done; done | tee gpurun_out/presets2.txt
for wc in 0 2 6; do
  TCGNN_WIN_COST=$wc timeout 300 python tools/quick.py --workload rmat-10m-200m --iters 3 --tag wincost$wc 2>&1 | tail -1
  TCGNN_WIN_COST=$wc timeout 300 python tools/quick.py --workload products-like-rmat --iters 3 --tag wincost$wc 2>&1 | tail -1
done | tee -a gpurun_out/presets2.txt
