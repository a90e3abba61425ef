#!/usr/bin/env bash
# Why the register-gather pipeline is slow: producer timeline, ablations, L1 no-allocate, one ncu --set full capture.
mkdir -p gpurun_out
export TCGNN_SPMM_TS=1
WL=reddit-like-rmat
echo "=== trace (TS, 3 teams)"
LD_LIBRARY_PATH=$PWD/variants/trace TCGNN_TRACE=$PWD/gpurun_out/trace_ts_$WL.bin timeout 200 python tools/quick.py --workload $WL --op spmm --iters 1 --tag trace 2>&1 | grep min_ms
python tools/trace.py gpurun_out/trace_ts_$WL.bin 2>&1 | tee gpurun_out/trace_ts_$WL.txt
echo "=== ablations (TS): 1 = no gathers, 2 = one MMA per window, 4 = no B tiles"
for a in 1 2 4; do
  LD_LIBRARY_PATH=$PWD/variants/trace TCGNN_ABLATE=$a timeout 200 python tools/quick.py --workload $WL --op spmm --iters 3 --tag ablate$a 2>&1 | grep min_ms
done
echo "=== L1::no_allocate"
LD_LIBRARY_PATH=$PWD/variants/noalloc timeout 200 python tools/ab.py --tag noalloc spmm:reddit-like-rmat spmm:reddit-like-uniform 2>&1 | grep "min_ms\|rror"
echo "=== ncu"
timeout 400 ncu --set full --clock-control none -k regex:spmm_tc_kernel -s 1 -c 1 -o /tmp/prof_ts -f \
    python tools/quick.py --workload $WL --op spmm --iters 1 > gpurun_out/ncu_ts.log 2>&1
python tools/ncu_summary.py /tmp/prof_ts.ncu-rep > gpurun_out/r02m_spmm_ts_${WL}_ncu_full.txt 2>&1
ncu -i /tmp/prof_ts.ncu-rep --page raw --csv > gpurun_out/r02m_spmm_ts_raw.csv 2>/dev/null
head -60 gpurun_out/r02m_spmm_ts_${WL}_ncu_full.txt | cut -c1-150
