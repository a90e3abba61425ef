#!/usr/bin/env bash
# multi-GPU check: sharding tests (incl. the world_size>1 NCCL test) and the strong-scaling bench line
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_gpu_sharding.py -m gpu -q --timeout 300 2>&1 | tail -5
for wl in ${WORKLOADS:-reddit-like-rmat}; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/bench_${wl}_n$N.json 2> gpurun_out/bench_${wl}_n$N.err
tail -c 2500 gpurun_out/bench_${wl}_n$N.json; tail -3 gpurun_out/bench_${wl}_n$N.err
done
