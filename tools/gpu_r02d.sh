#!/usr/bin/env bash
mkdir -p gpurun_out
for cfg in "1 0" "2 0" "2 1" "3 0"; do
  set -- $cfg
  TCGNN_SDDMM_TEAM=$1 TCGNN_SDDMM_DBG=$2 timeout 600 python tools/stress_sddmm.py products-like-rmat 256 24 2>&1 | grep -v Warn | tail -12
done | tee gpurun_out/stress_sddmm.txt
echo "=== spmm publish / team A-B"
for wl in reddit-like-rmat reddit-like-uniform products-like-rmat; do
for cfg in "1 1" "1 0" "2 1" "2 0"; do
  set -- $cfg
  TCGNN_SPMM_TEAM=$1 TCGNN_SPMM_LAG=$2 timeout 300 python tools/quick.py --workload $wl --op spmm --iters 3 --tag team$1lag$2 2>&1 | tail -1
  TCGNN_SPMM_TEAM=$1 TCGNN_SPMM_LAG=$2 timeout 300 python tools/quick.py --workload $wl --op wspmm_tile --iters 3 --tag team$1lag$2 2>&1 | tail -1
done; done | tee gpurun_out/timings_d.txt
for t in 1 2 3; do TCGNN_SDDMM_TEAM=$t timeout 300 python tools/quick.py --workload reddit-like-uniform --op sddmm --iters 3 --tag sddmm_team$t 2>&1 | tail -1; TCGNN_SDDMM_TEAM=$t timeout 300 python tools/quick.py --workload products-like-rmat --op sddmm --iters 3 --tag sddmm_team$t 2>&1 | tail -1; done | tee -a gpurun_out/timings_d.txt
echo "=== parity with eager publish"; timeout 900 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_fused_ops.py -m gpu -q --timeout 600 2>&1 | tail -3
