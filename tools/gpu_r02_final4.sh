#!/usr/bin/env bash
# Last check of the final tree (SpMM kernel as validated in r02n + SDDMM plain path as validated in final3): parity of
# the touched operators and the default bench line.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_sddmm.py tests/test_gpu_fused_ops.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_final4.json 2> gpurun_out/bench_final4.err; echo "rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/bench_final4.json').read()); print(d['ms_per_step'], d['min_ms'], d['parity']['bit_exact'], d['e2e']['ms_per_step'], {k:v.get('ms_per_step') for k,v in d['variants'].items()})"
