#!/usr/bin/env bash
N=8
mkdir -p gpurun_out
run() {
  name=$1; shift
  echo "=== bench $name (N=$N)"
  env "$@" timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 20 --warmup 5 $BARGS > gpurun_out/bench_${name}_n$N.json 2> gpurun_out/bench_${name}_n$N.err
  echo "rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${name}_n$N.json").read())
    print({k: d.get(k) for k in ("ms_per_step", "min_ms", "value", "parity", "exchange")})
    print("e2e", (d.get("e2e") or {}).get("ms_per_step"), "variants", {k: (v.get("ms_per_step"), v.get("parity", {}).get("max_abs_diff"), v.get("exchange")) for k, v in (d.get("variants") or {}).items()})
except Exception as exc:
    print("no json:", exc)
PY
  grep -v "Warn\|sparse_csr\|OMP_NUM\|\*\*\*\*\|NCCL version" gpurun_out/bench_${name}_n$N.err | tail -6 | cut -c1-300
}
BARGS="" run overlap_graph TCGNN_EXCHANGE=auto
