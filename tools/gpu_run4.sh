#!/usr/bin/env bash
# re-entry pass: every GPU parity file (one process each: a trapped kernel poisons the CUDA context),
# smoke, ablation timings of the SpMM pipeline, bench (rmat + uniform), ncu launch list + full capture.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in test_gpu_umma_layouts test_gpu_sgt test_gpu_spmm test_gpu_sddmm test_gpu_vs_reference test_gpu_sharding test_gpu_layers; do
  echo "=== $f"
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 300 -x 2>&1 | tail -40 > gpurun_out/$f.log
  tail -3 gpurun_out/$f.log
done
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "=== ablations (1 = no gathers, 2 = no MMAs, 4 = no B build)"
for wl in reddit-like-uniform reddit-like-rmat; do
for ab in 0 1 2 4 7; do
  TCGNN_ABLATE=$ab timeout 300 python tools/quick.py --workload $wl --iters 3 2>&1 | tail -1
done; done | tee gpurun_out/ablate.txt
for op in sddmm wspmm; do timeout 300 python tools/quick.py --workload reddit-like-uniform --op $op --iters 3 2>&1 | tail -1; done | tee -a gpurun_out/ablate.txt
timeout 300 python tools/quick.py --workload products-like-rmat --iters 3 2>&1 | tail -1 | tee -a gpurun_out/ablate.txt
timeout 300 python tools/quick.py --workload citeseer-like --iters 20 2>&1 | tail -1 | tee -a gpurun_out/ablate.txt
echo "=== bench ours"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -c 3000 gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
echo "=== bench uniform"; timeout 900 python bench.py --steps 10 --warmup 3 --workload reddit-like-uniform --no-cpu-baseline > gpurun_out/bench_uniform.json 2> gpurun_out/bench_uniform.err; tail -c 3000 gpurun_out/bench_uniform.json; tail -5 gpurun_out/bench_uniform.err
echo "=== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'spmm_tc_kernel|tf32_round' -s 6 -c 2 -o gpurun_out/prof_spmm_uniform -f python bench.py --steps 2 --warmup 3 --workload reddit-like-uniform --no-cpu-baseline --no-reference-gpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
