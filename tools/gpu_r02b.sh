#!/usr/bin/env bash
# Round 2, second single-GPU pass: parity of the changed kernels (producer teams, SDDMM tile-order output), timings
# with one vs two gathering warps per stage, SDDMM determinism diagnostic, L2 gather sweep with per-warp random rows.
mkdir -p gpurun_out
for f in test_gpu_spmm test_gpu_sddmm test_gpu_fused_ops test_gpu_sharding test_gpu_layers test_gpu_fullsize; do
  echo "=== $f"
  timeout 1200 python -m pytest tests/$f.py -m gpu -q --timeout 900 2>&1 | grep -v Warn | tail -60 > gpurun_out/$f.log
  tail -3 gpurun_out/$f.log
done
echo "=== diag"; timeout 600 python tools/diag_sddmm.py rmat-10m-200m 2>&1 | tail -6 | tee gpurun_out/diag_sddmm.txt
echo "=== timings"
for wl in reddit-like-rmat reddit-like-uniform products-like-rmat rmat-10m-200m; do
for op in spmm sddmm wspmm_tile agnn spmm_host; do
  timeout 300 python tools/quick.py --workload $wl --op $op --iters 3 --tag team2 2>&1 | tail -1
done
TCGNN_SPMM_TEAM=1 timeout 300 python tools/quick.py --workload $wl --op spmm --iters 3 --tag team1 2>&1 | tail -1
done | tee gpurun_out/timings_b.txt
echo "=== l2 gather ceiling / gather4 A/B"
timeout 600 python tools/l2_gather_bench.py --out gpurun_out/l2_gather 2>&1 | tail -3
echo "=== bench ours"; timeout 900 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -c 1500 gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
