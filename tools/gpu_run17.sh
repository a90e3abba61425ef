#!/usr/bin/env bash
mkdir -p gpurun_out
for ab in 0 256 512; do
  TCGNN_WIN_COST=0 TCGNN_ABLATE=$ab timeout 300 python tools/quick.py --workload rmat-10m-200m --iters 2 --tag ablate$ab 2>&1 | tail -1
done | tee gpurun_out/epi.txt
TCGNN_WIN_COST=0 TCGNN_TRACE_CTA=127 TCGNN_TRACE=gpurun_out/trace_cta127.bin timeout 300 python tools/quick.py --workload rmat-10m-200m --iters 1 --tag cta127 2>&1 | tail -1
python tools/trace.py gpurun_out/trace_cta127.bin 50 500 | tail -12 | tee -a gpurun_out/epi.txt
TCGNN_PRESET=1 timeout 300 python tools/quick.py --workload reddit-like-uniform --iters 2 --tag preset1 2>&1 | tail -3
