#!/usr/bin/env bash
# Final single-GPU pass: what the driver runs (pytest -m gpu in one process, smoke, the default bench line) + operator timings.
mkdir -p gpurun_out
echo "=== pytest -m gpu"; timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_gpu_final3.txt 2>&1; tail -3 gpurun_out/pytest_gpu_final3.txt
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "=== bench ours"; timeout 600 python bench.py > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_final3.json").read())
print({k: d.get(k) for k in ("ms_per_step", "min_ms", "incomplete", "gpu_launches")}, "bit_exact", d["parity"]["bit_exact"], "e2e", d["e2e"]["ms_per_step"])
print("variants", {k: (v.get("ms_per_step"), (v.get("reference_gpu") or {}).get("speedup_device")) for k, v in d["variants"].items()})
PY
echo "=== timings"
timeout 300 python tools/ab.py --tag final3 wspmm:reddit-like-rmat sddmm:reddit-like-rmat agnn:reddit-like-rmat sddmm:reddit-like-uniform agnn:reddit-like-uniform sddmm:products-like-rmat agnn:products-like-rmat 2>&1 | grep "min_ms\|rror" | tee gpurun_out/timings_final3.txt
