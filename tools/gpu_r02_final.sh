#!/usr/bin/env bash
# Final single-GPU pass of round 2: what the driver runs (pytest -m gpu in one process, smoke, bench both arms), plus
# the AGNN bench line, the reorder report and the ncu launch list of the default bench command.
mkdir -p gpurun_out
echo "=== pytest -m gpu (one process, as the driver runs it)"; timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | grep -v Warn | tail -4 | tee gpurun_out/pytest_gpu_final.txt
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench ours"; timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo rc=$?; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_final.json"))
print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches", "parity")})
print("e2e", d["e2e"]["ms_per_step"], d["e2e"].get("max_rel_diff_vs_resident"))
print("roofline", {k: v for k, v in d["roofline"].items() if k not in ("note",)})
print("variants", {k: {kk: v.get(kk) for kk in ("ms_per_step", "reference_gpu", "nnz_per_tc_block")} for k, v in d.get("variants", {}).items()})
print("cpu", d.get("cpu_baseline", {}).get("value"), d.get("cpu_baseline_dgl", {}).get("train_ms_per_epoch"))
PY
tail -3 gpurun_out/bench_final.err | cut -c1-200
echo "=== bench agnn products"; timeout 900 python bench.py --workload products-like-rmat --op agnn --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_agnn_products.json 2> gpurun_out/bench_agnn_products.err; python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_agnn_products.json"))
print({k: d.get(k) for k in ("value", "ms_per_step", "parity")}, "e2e", d["e2e"]["ms_per_step"], d["e2e"]["api"])
PY
echo "=== reorder report"; timeout 900 python tools/reorder_report.py --out gpurun_out/reorder_tc_blocks.csv 2>&1 | grep -v Warn | tail -14
echo "=== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmm|tf32_round|zero_partial|sddmm|unpermute|permute|gather_rows|wait_flag' -c 60 --csv --log-file gpurun_out/r02j_launches_bench_default.csv \
   python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/ncu_launch.log 2>&1
grep -c spmm_tc gpurun_out/r02j_launches_bench_default.csv
echo "=== bench reference arm"; TCGNN_REF_BUDGET_S=60 timeout 900 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err; head -c 600 gpurun_out/bench_ref_final.json; echo
