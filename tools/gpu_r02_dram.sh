#!/usr/bin/env bash
# DRAM bytes of the final SpMM kernel (no L2 hint on the gathers): ncu, two metrics, one launch per workload.
mkdir -p gpurun_out
for wl in reddit-like-rmat reddit-like-uniform; do
  timeout 60 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none \
     -k regex:spmm_tc_kernel -s 1 -c 1 --csv --log-file gpurun_out/r02o_dram_spmm_$wl.csv python tools/quick.py --workload $wl --op spmm --iters 1 > /dev/null 2>&1
  grep -v "^==" gpurun_out/r02o_dram_spmm_$wl.csv | awk -F'","' '{print $(NF-2), $(NF-1), $NF}' | tail -4
done
