#!/usr/bin/env bash
# first bring-up run: each test file in its own process (a trapped kernel poisons the CUDA context)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in test_gpu_umma_layouts test_gpu_sgt test_gpu_spmm test_gpu_sddmm test_gpu_vs_reference; do
  echo "=== $f" 
  timeout 600 python -m pytest tests/$f.py -m gpu -q --timeout 120 -x 2>&1 | tail -40 > gpurun_out/$f.log
  tail -5 gpurun_out/$f.log
done
cat gpurun_out/umma_probe.txt
