#!/usr/bin/env bash
mkdir -p gpurun_out
for f in test_gpu_spmm test_gpu_sddmm test_gpu_vs_reference test_gpu_layers test_gpu_fullsize; do
  echo "=== $f"
  timeout 900 python -m pytest tests/$f.py -m gpu -q --timeout 600 -x 2>&1 | grep -v Warn | tail -3
done
for wl in reddit-like-rmat reddit-like-uniform products-like-rmat; do
for op in spmm sddmm wspmm; do
  timeout 300 python tools/quick.py --workload $wl --op $op --iters 3 2>&1 | tail -1
done; done | tee gpurun_out/ops.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmm|sddmm|permute|tf32_round|zero_partial|unpermute' -c 30 --csv --log-file gpurun_out/launches_ops.csv python tools/quick.py --workload reddit-like-uniform --op wspmm --iters 2 > /dev/null 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'spmm|sddmm|permute|tf32_round|zero_partial|unpermute' -c 30 --csv --log-file gpurun_out/launches_sddmm.csv python tools/quick.py --workload reddit-like-uniform --op sddmm --iters 2 > /dev/null 2>&1
python - <<'PY'
import csv
from collections import defaultdict
for f in ("gpurun_out/launches_ops.csv", "gpurun_out/launches_sddmm.csv"):
    rows=[r for r in csv.reader(open(f)) if len(r)>10 and r[0].isdigit()]
    d=defaultdict(list)
    for r in rows: d[r[4][:70]].append(float(r[-1])/1e3)
    print(f)
    for k,v in d.items(): print(f"   {k:72s} n={len(v):3d} avg_us={sum(v)/len(v):10.1f}")
PY
