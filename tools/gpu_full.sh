#!/usr/bin/env bash
# Full single-GPU pass (what the driver runs at round end, plus profiles): every GPU parity file (one process each:
# a trapped kernel poisons the CUDA context), smoke, the reference's caller, bench (both arms), ncu launch list and
# full captures.  Outputs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in test_gpu_umma_layouts test_gpu_sgt test_gpu_spmm test_gpu_sddmm test_gpu_vs_reference test_gpu_sharding test_gpu_layers test_gpu_fullsize; do
  echo "=== $f"
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 600 -x 2>&1 | grep -v Warn | tail -40 > gpurun_out/$f.log
  tail -2 gpurun_out/$f.log
done
echo "=== pytest -m gpu (one process, as the driver runs it)"; timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== main_tcgnn (single kernel + 3 GCN epochs on a citeseer-sized synthetic graph)"
(cd tc-gnn_atc23_b200 && timeout 300 python main_tcgnn.py --dataset citeseer --dim 16 --hidden 16 --classes 6 --single_kernel 2>&1 | tail -4; timeout 300 python main_tcgnn.py --dataset citeseer --dim 64 --hidden 16 --classes 6 --epochs 3 2>&1 | tail -4)
echo "=== timings"
for wl in reddit-like-rmat reddit-like-uniform products-like-rmat rmat-10m-200m citeseer-like; do
for op in spmm sddmm wspmm; do
  timeout 300 python tools/quick.py --workload $wl --op $op --iters 3 2>&1 | tail -1
done; done | tee gpurun_out/timings.txt
echo "=== bench ours"; timeout 900 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -c 3800 gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
echo "=== bench uniform"; timeout 900 python bench.py --workload reddit-like-uniform --no-cpu-baseline > gpurun_out/bench_uniform.json 2> gpurun_out/bench_uniform.err; tail -c 1200 gpurun_out/bench_uniform.json; tail -3 gpurun_out/bench_uniform.err
echo "=== bench agnn products"; timeout 900 python bench.py --workload products-like-rmat --op agnn --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_agnn.json 2> gpurun_out/bench_agnn.err; tail -c 1200 gpurun_out/bench_agnn.json; tail -3 gpurun_out/bench_agnn.err
echo "=== bench reference arm"; timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 1500 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
echo "=== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmm|tf32_round|zero_partial' -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log | cut -c1-200
echo "=== ncu full"
for wl in reddit-like-rmat reddit-like-uniform; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'spmm_tc_kernel' -s 3 -c 1 -o gpurun_out/prof_spmm_$wl -f python bench.py --steps 2 --warmup 3 --workload $wl --no-cpu-baseline --no-reference-gpu > gpurun_out/ncu_full_$wl.log 2>&1
tail -1 gpurun_out/ncu_full_$wl.log
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sddmm_tc_kernel' -s 1 -c 1 -o gpurun_out/prof_sddmm_uniform -f python tools/quick.py --workload reddit-like-uniform --op sddmm --iters 2 > gpurun_out/ncu_full_sddmm.log 2>&1
tail -1 gpurun_out/ncu_full_sddmm.log
