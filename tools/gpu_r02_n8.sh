#!/usr/bin/env bash
# 8-GPU pass: the driver's scaling command (default line incl. the R-MAT 10M variant) with the overlapped exchange,
# then the round-1 gather-then-compute exchange for comparison.
N=${1:-8}
mkdir -p gpurun_out
run() {
  name=$1; shift
  echo "=== bench $name (N=$N)"
  env "$@" timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 20 --warmup 5 $BARGS > gpurun_out/bench_${name}_n$N.json 2> gpurun_out/bench_${name}_n$N.err
  echo "rc=$?"; python - <<PY
import json
try:
    txt = open("gpurun_out/bench_${name}_n$N.json").read()
    d = json.loads(txt[txt.index("{"):])
    print({k: d.get(k) for k in ("ms_per_step", "min_ms", "value", "parity", "exchange")})
    print("e2e", (d.get("e2e") or {}).get("ms_per_step"), "variants", {k: (v.get("ms_per_step"), v.get("parity", {}).get("max_abs_diff"), v.get("exchange")) for k, v in (d.get("variants") or {}).items()})
except Exception as exc:
    print("no json:", exc)
PY
  grep -v "Warn\|sparse_csr\|OMP_NUM\|\*\*\*\*" gpurun_out/bench_${name}_n$N.err | tail -6 | cut -c1-300
}
BARGS="" run overlap TCGNN_EXCHANGE=auto
BARGS="--no-variants --no-e2e" run gather_p2p2 TCGNN_EXCHANGE=p2p2
BARGS="--workload rmat-10m-200m --steps 10 --warmup 3 --no-e2e" run c5_gather TCGNN_EXCHANGE=p2p2
