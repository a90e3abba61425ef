#!/usr/bin/env bash
# Final single-GPU pass after the gather-loop trimming: what the driver runs (pytest -m gpu in one process, smoke,
# the default bench line) + the AGNN products line.
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu_final2.txt 2>&1; tail -3 gpurun_out/pytest_gpu_final2.txt
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench ours"; timeout 600 python bench.py > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_final2.json").read())
print({k: d.get(k) for k in ("ms_per_step", "min_ms", "incomplete", "gpu_launches")}, "bit_exact", d["parity"]["bit_exact"], "e2e", d["e2e"]["ms_per_step"])
print("roofline", {k: d["roofline"].get(k) for k in ("frac", "dram_frac", "kernel_ms")}, d["roofline"]["l2_gather"]["frac"], d["roofline"]["l2_gather"].get("frac_same_working_set"))
print("variants", {k: (v.get("ms_per_step"), (v.get("reference_gpu") or {}).get("speedup_device")) for k, v in d["variants"].items()})
PY
echo "=== bench agnn products"; timeout 600 python bench.py --workload products-like-rmat --op agnn > gpurun_out/bench_agnn_products2.json 2> gpurun_out/bench_agnn_products2.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_agnn_products2.json').read()); print(d['ms_per_step'], d['min_ms'], d['parity']['bit_exact'], d['e2e']['ms_per_step'])"
