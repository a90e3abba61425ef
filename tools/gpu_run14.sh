#!/usr/bin/env bash
mkdir -p gpurun_out
for wl in rmat-10m-200m products-like-rmat reddit-like-rmat citeseer-like; do
TCGNN_TRACE=gpurun_out/trace_$wl.bin timeout 300 python tools/quick.py --workload $wl --iters 1 --tag trace 2>&1 | tail -1
python tools/trace.py gpurun_out/trace_$wl.bin | tail -8
done | tee gpurun_out/percta.txt
