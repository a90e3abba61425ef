#!/usr/bin/env bash
# round-1 GPU pass: parity tests (one process per file: a trapped kernel poisons the CUDA context),
# smoke, bench (ours + reference arm), ncu launch list and one full capture of the SpMM kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in test_gpu_spmm test_gpu_sddmm test_gpu_sgt test_gpu_vs_reference test_gpu_sharding test_gpu_layers test_gpu_umma_layouts; do
  echo "=== $f"
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 300 -x 2>&1 | tail -40 > gpurun_out/$f.log
  tail -3 gpurun_out/$f.log
done
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "=== bench ours"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -c 3000 gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
echo "=== bench uniform"; timeout 900 python bench.py --steps 10 --warmup 3 --workload reddit-like-uniform --no-cpu-baseline > gpurun_out/bench_uniform.json 2> gpurun_out/bench_uniform.err; tail -c 3000 gpurun_out/bench_uniform.json; tail -5 gpurun_out/bench_uniform.err
echo "=== bench reference"; timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 2000 gpurun_out/bench_ref.json; tail -5 gpurun_out/bench_ref.err
echo "=== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmm|sddmm|zero_partial|permute' -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmm_tc_kernel -s 3 -c 1 -o gpurun_out/prof_spmm -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
