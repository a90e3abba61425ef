#!/usr/bin/env bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharding.py -m gpu -q --timeout 300 -x 2>&1 | tail -30
for mode in nccl p2p auto; do
TCGNN_EXCHANGE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/dbg_$mode.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$mode', 'ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'launches', d['gpu_launches'])
"
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/dbg_$mode.err | tail -5 | cut -c1-300
done
