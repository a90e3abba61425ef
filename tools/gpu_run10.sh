#!/usr/bin/env bash
mkdir -p gpurun_out
for ps in 0 3; do
  echo "=== test_gpu_spmm preset $ps"
  TCGNN_PRESET=$ps timeout 900 python -m pytest tests/test_gpu_spmm.py -m gpu -q --timeout 300 -x 2>&1 | tail -3
done
for wl in reddit-like-uniform reddit-like-rmat; do
for ps in 0 1 2 3 4 5 6; do
  TCGNN_PRESET=$ps timeout 300 python tools/quick.py --workload $wl --iters 3 --tag preset$ps 2>&1 | tail -1
done; done | tee gpurun_out/presets.txt
for ps in 0 3; do for ab in 1 5 13; do
  TCGNN_PRESET=$ps TCGNN_ABLATE=$ab timeout 300 python tools/quick.py --workload reddit-like-uniform --iters 3 --tag p${ps}a$ab 2>&1 | tail -1
done; done | tee -a gpurun_out/presets.txt
for ps in 0 2 3; do
TCGNN_PRESET=$ps timeout 300 python tools/quick.py --workload products-like-rmat --iters 3 --tag preset$ps 2>&1 | tail -1 
TCGNN_PRESET=$ps timeout 300 python tools/quick.py --workload reddit-like-uniform --op wspmm --iters 3 --tag preset$ps 2>&1 | tail -1 
done | tee -a gpurun_out/presets.txt
