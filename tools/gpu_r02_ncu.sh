#!/usr/bin/env bash
# ncu --set full captures of the dominant kernels (one launch each), summarised on the box (the reports themselves
# are too large to bring back), plus the launch list of the default bench command.
mkdir -p gpurun_out
cap() {  # name workload op kernel-regex skip
  timeout 600 ncu --set full --clock-control none -k regex:$4 -s $5 -c 1 -o /tmp/prof_$1 -f \
      python tools/quick.py --workload $2 --op $3 --iters 1 > gpurun_out/ncu_$1.log 2>&1
  python tools/ncu_summary.py /tmp/prof_$1.ncu-rep > gpurun_out/r02i_$1_ncu_full.txt 2>&1
  grep -E "gpu__time_duration.sum|dram__bytes_read.sum |dram__bytes_write.sum |lts__t_sector_hit_rate|sm__pipe_tensor_cycles_active.avg" gpurun_out/r02i_$1_ncu_full.txt | head -5
  rm -f /tmp/prof_$1.ncu-rep
}
cap spmm_reddit-like-rmat_D128 reddit-like-rmat spmm spmm_tc_kernel 1
cap spmm_reddit-like-uniform_D128 reddit-like-uniform spmm spmm_tc_kernel 1
cap spmm_products-like-rmat_D256 products-like-rmat spmm spmm_tc_kernel 1
cap spmm_rmat-10m-200m_D256 rmat-10m-200m spmm spmm_tc_kernel 1
cap wspmm_reddit-like-uniform_D128 reddit-like-uniform wspmm_tile spmm_tc_kernel 2
cap sddmm_reddit-like-uniform_D128 reddit-like-uniform sddmm sddmm_tc_kernel 1
cap sddmm_products-like-rmat_D256 products-like-rmat sddmm sddmm_tc_kernel 1
echo "=== launch list of the default bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02i_launches_bench_default.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-200
