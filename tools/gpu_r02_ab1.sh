#!/usr/bin/env bash
# Where a producer's time goes per stage (trace build) + wait-policy variants of the mbarrier waits.
mkdir -p gpurun_out
for wl in reddit-like-rmat reddit-like-uniform; do
  LD_LIBRARY_PATH=$PWD/variants/trace TCGNN_TRACE=$PWD/gpurun_out/trace_$wl.bin timeout 200 python tools/quick.py --workload $wl --op spmm --iters 1 --tag trace 2>&1 | grep min_ms
  python tools/trace.py gpurun_out/trace_$wl.bin 2>&1 | tee gpurun_out/trace_$wl.txt
done
ITEMS="spmm:reddit-like-rmat wspmm:reddit-like-rmat sddmm:reddit-like-rmat spmm:reddit-like-uniform sddmm:reddit-like-uniform"
timeout 200 python tools/ab.py --tag base $ITEMS 2>&1 | grep min_ms
for v in hint20 backoff hintbackoff; do
  LD_LIBRARY_PATH=$PWD/variants/$v timeout 200 python tools/ab.py --tag $v $ITEMS 2>&1 | grep min_ms
done
