#!/usr/bin/env bash
# Round-2 check after restructuring bench.py (guarded optional sections) + where the fused AGNN time goes on products.
mkdir -p gpurun_out
echo "=== bench default"
timeout 600 python bench.py > gpurun_out/bench_check3.json 2> gpurun_out/bench_check3.err; echo "rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_check3.json").read())
print({k: d.get(k) for k in ("ms_per_step", "incomplete")}, "e2e", (d.get("e2e") or {}).get("ms_per_step"),
      "variants", {k: v.get("ms_per_step") for k, v in (d.get("variants") or {}).items()}, "cpu", (d.get("cpu_baseline") or {}).get("value"))
PY
echo "=== products AGNN parts"
for op in sddmm wspmm_tile agnn spmm; do
  timeout 300 python tools/quick.py --workload products-like-rmat --op $op --iters 8 --tag final 2>&1 | grep min_ms
done
