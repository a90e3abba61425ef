#!/usr/bin/env bash
mkdir -p gpurun_out
for ab in 0 13 5 1; do
  TCGNN_ABLATE=$ab TCGNN_TRACE=gpurun_out/trace_a$ab.bin timeout 300 python tools/quick.py --workload reddit-like-uniform --iters 2 --tag ablate$ab 2>&1 | tail -1
  echo "--- trace ablate=$ab"; python tools/trace.py gpurun_out/trace_a$ab.bin
done 2>&1 | tee gpurun_out/trace.txt
