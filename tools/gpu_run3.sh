#!/usr/bin/env bash
# iteration pass: parity of the changed kernels, bench (rmat + uniform), ncu full of spmm
mkdir -p gpurun_out
for f in test_gpu_spmm test_gpu_sddmm test_gpu_vs_reference test_gpu_sharding test_gpu_layers; do
  echo "=== $f"
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 300 -x 2>&1 | tail -40 > gpurun_out/$f.log
  tail -3 gpurun_out/$f.log
done
echo "=== bench ours"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline ${BENCH_EXTRA:-} > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; python - <<'PY'
import json
for f in ["gpurun_out/bench_ours.json"]:
    try:
        d=json.load(open(f)); print({k:d[k] for k in ("value","ms_per_step","min_ms")}, d["roofline"]["frac"], d["e2e"]["ms_per_step"], d.get("reference_gpu",{}).get("ms_per_step"), d.get("reference_gpu",{}).get("max_abs_diff_vs_ours"))
    except Exception as e: print("ERR", e)
PY
tail -5 gpurun_out/bench_ours.err
echo "=== bench uniform"; timeout 900 python bench.py --steps 10 --warmup 3 --workload reddit-like-uniform --no-cpu-baseline ${BENCH_EXTRA:-} > gpurun_out/bench_uniform.json 2> gpurun_out/bench_uniform.err; python - <<'PY'
import json
for f in ["gpurun_out/bench_uniform.json"]:
    try:
        d=json.load(open(f)); print({k:d[k] for k in ("value","ms_per_step","min_ms")}, d["roofline"]["frac"], d["e2e"]["ms_per_step"], d.get("reference_gpu",{}).get("ms_per_step"), d.get("reference_gpu",{}).get("max_abs_diff_vs_ours"))
    except Exception as e: print("ERR", e)
PY
tail -5 gpurun_out/bench_uniform.err
if [ -n "${AGNN:-}" ]; then
echo "=== bench agnn products"; timeout 900 python bench.py --steps 5 --warmup 3 --workload products-like-rmat --op agnn --no-cpu-baseline --no-reference-gpu > gpurun_out/bench_agnn.json 2> gpurun_out/bench_agnn.err; tail -c 1500 gpurun_out/bench_agnn.json; tail -5 gpurun_out/bench_agnn.err
fi
echo "=== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'spmm_tc_kernel|tf32_round' -s 6 -c 2 -o gpurun_out/prof_spmm -f python bench.py --steps 2 --warmup 3 --workload ${NCU_WORKLOAD:-reddit-like-uniform} --no-cpu-baseline --no-reference-gpu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
