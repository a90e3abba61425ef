#!/usr/bin/env bash
mkdir -p gpurun_out
for cta in 127 60; do
TCGNN_TRACE_CTA=$cta TCGNN_TRACE=gpurun_out/trace_cta$cta.bin timeout 300 python tools/quick.py --workload rmat-10m-200m --iters 1 --tag cta$cta 2>&1 | tail -1
python tools/trace.py gpurun_out/trace_cta$cta.bin 50 500 | grep -v "^   CTA\|least\|per-CTA"
done | tee gpurun_out/percta2.txt
