#!/usr/bin/env bash
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/exchange_probe2.py reddit-like-rmat > gpurun_out/exchange_probe2_n$N.log 2>&1
grep -v "Warn\|OMP_NUM\|\*\*\*\*\|NCCL version" gpurun_out/exchange_probe2_n$N.log | grep -B2 -A12 "Traceback" | head -60
grep "max over ranks\|panel rows" gpurun_out/exchange_probe2_n$N.log | tee gpurun_out/exchange_probe2_n$N.txt
