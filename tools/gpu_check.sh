#!/usr/bin/env bash
# quick re-validation after host-side changes: the driver's own commands
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --no-reference-gpu > gpurun_out/bench_check.json 2> gpurun_out/bench_check.err; python -c "
import json; d=json.load(open('gpurun_out/bench_check.json')); print({k: d[k] for k in ('value','ms_per_step','gpu_launches','dtype','steps','warmup')}, d['e2e']['ms_per_step'], d['roofline']['frac'], d['roofline']['l2_gather']['frac'], d['cpu_baseline']['value'])"; tail -2 gpurun_out/bench_check.err
TCGNN_REF_BUDGET_S=40 timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_check.json 2> gpurun_out/bench_ref_check.err; python -c "
import json; d=json.load(open('gpurun_out/bench_ref_check.json')); print({k: d[k] for k in ('impl','value','ms_per_step','dtype','steps','warmup')})"; tail -2 gpurun_out/bench_ref_check.err
