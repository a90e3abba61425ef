#!/usr/bin/env bash
mkdir -p gpurun_out
{
TCGNN_SDDMM_TEAM=2 timeout 900 python tools/stress_sddmm.py products-like-rmat 256 60 2>&1 | grep -v Warn | tail -8
TCGNN_SDDMM_TEAM=1 timeout 900 python tools/stress_sddmm.py rmat-10m-200m 256 20 2>&1 | grep -v Warn | tail -8
timeout 900 python tools/stress_sddmm.py products-like-rmat 256 40 wspmm 2>&1 | grep -v Warn | tail -8
timeout 900 python tools/stress_sddmm.py products-like-rmat 256 20 spmm 2>&1 | grep -v Warn | tail -8
} | tee gpurun_out/stress_after_fix.txt
for f in test_gpu_spmm test_gpu_sddmm test_gpu_fused_ops test_gpu_fullsize; do
  echo "=== $f"
  timeout 1200 python -m pytest tests/$f.py -m gpu -q --timeout 900 2>&1 | grep -v Warn | tail -60 > gpurun_out/$f.log
  tail -3 gpurun_out/$f.log
done
