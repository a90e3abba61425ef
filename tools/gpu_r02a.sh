#!/usr/bin/env bash
# Round 2, first single-GPU pass: every GPU parity file (one process each: a trapped kernel poisons the CUDA
# context), the L2 gather ceiling + gather4 A/B, op timings, the default bench line.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in test_gpu_umma_layouts test_gpu_sgt test_gpu_spmm test_gpu_sddmm test_gpu_fused_ops test_gpu_vs_reference test_gpu_reference_driver test_gpu_sharding test_gpu_layers test_gpu_fullsize; do
  echo "=== $f"
  timeout 1200 python -m pytest tests/$f.py -m gpu -q --timeout 900 2>&1 | grep -v Warn | tail -60 > gpurun_out/$f.log
  tail -3 gpurun_out/$f.log
done
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== l2 gather ceiling / gather4 A/B"
timeout 600 python tools/l2_gather_bench.py --out gpurun_out/l2_gather 2>&1 | tail -3
echo "=== timings"
for wl in reddit-like-rmat reddit-like-uniform products-like-rmat; do
for op in spmm sddmm wspmm wspmm_tile agnn agnn3 spmm_host; do
  timeout 300 python tools/quick.py --workload $wl --op $op --iters 3 2>&1 | tail -1
done; done | tee gpurun_out/timings.txt
echo "=== bench ours"; timeout 900 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -c 4500 gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
