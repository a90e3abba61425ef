#!/usr/bin/env bash
mkdir -p gpurun_out
echo "=== test_gpu_fullsize"
timeout 900 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --timeout 600 -x 2>&1 | tail -15
echo "=== bench ours (default)"; timeout 900 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; tail -c 3500 gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
echo "=== bench reference arm"; timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -c 2500 gpurun_out/bench_ref.json; tail -3 gpurun_out/bench_ref.err
echo "=== bench rmat-10m-200m N=1"; timeout 900 python bench.py --workload rmat-10m-200m --steps 5 --warmup 3 --no-reference-gpu > gpurun_out/bench_rmat10m_n1.json 2> gpurun_out/bench_rmat10m_n1.err; tail -c 2500 gpurun_out/bench_rmat10m_n1.json; tail -3 gpurun_out/bench_rmat10m_n1.err
echo "=== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmm|tf32_round|zero_partial' -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-300
echo "=== ncu full rmat"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'spmm_tc_kernel' -s 3 -c 1 -o gpurun_out/prof_spmm_rmat -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
