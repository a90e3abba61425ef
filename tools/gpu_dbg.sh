#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharding.py -m gpu -q --timeout 300 -x 2>&1 | tail -40
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | grep -v "^\*\*\*\|OMP_NUM" | tail -30 | cut -c1-1500
echo "exit: $?"
