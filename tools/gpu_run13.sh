#!/usr/bin/env bash
mkdir -p gpurun_out
for ab in 0 1 5; do
  TCGNN_ABLATE=$ab timeout 300 python tools/quick.py --workload rmat-10m-200m --iters 2 --tag ablate$ab 2>&1 | tail -1
done | tee gpurun_out/rmat10m.txt
for tune in 0 16 96; do
  TCGNN_TUNE=$tune timeout 300 python tools/quick.py --workload rmat-10m-200m --iters 2 --tag tune$tune 2>&1 | tail -1
done | tee -a gpurun_out/rmat10m.txt
timeout 300 python tools/quick.py --workload rmat-10m-200m --dim 128 --iters 2 --tag D128 2>&1 | tail -1 | tee -a gpurun_out/rmat10m.txt
timeout 300 python tools/quick.py --workload rmat-10m-200m --dim 32 --iters 2 --tag D32 2>&1 | tail -1 | tee -a gpurun_out/rmat10m.txt
TCGNN_TRACE=gpurun_out/trace_rmat10m.bin timeout 300 python tools/quick.py --workload rmat-10m-200m --iters 1 --tag trace 2>&1 | tail -1
python tools/trace.py gpurun_out/trace_rmat10m.bin | tee -a gpurun_out/rmat10m.txt
