#!/usr/bin/env bash
# last pass of the round: the driver's test command on the final tree, AGNN bench line, timings of every op
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -1
timeout 600 python bench.py --workload products-like-rmat --op agnn --steps 5 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_agnn.json 2> gpurun_out/bench_agnn.err; python -c "
import json; d=json.load(open('gpurun_out/bench_agnn.json')); print('agnn products', d['ms_per_step'], 'ms', round(d['value']/1e9,2), 'Gedges/s e2e', d['e2e']['ms_per_step'])"; tail -2 gpurun_out/bench_agnn.err
for wl in reddit-like-rmat reddit-like-uniform products-like-rmat rmat-10m-200m; do
for op in spmm sddmm wspmm; do
  timeout 300 python tools/quick.py --workload $wl --op $op --iters 3 2>&1 | tail -1
done; done | tee gpurun_out/timings_final.txt
