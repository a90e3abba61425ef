#!/usr/bin/env bash
mkdir -p gpurun_out
run() {
  name=$1; n=$2; shift; shift
  echo "=== bench $name (N=$n) $@"
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 20 --warmup 5 $BARGS > gpurun_out/bench_${name}_n$n.json 2> gpurun_out/bench_${name}_n$n.err
  echo "rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_${name}_n$n.json").read())
    print({k: d.get(k) for k in ("ms_per_step", "min_ms", "partition")}, "bit_exact", d["parity"]["bit_exact"])
    print("e2e", (d.get("e2e") or {}).get("ms_per_step"), "variants", {k: (v.get("ms_per_step"), v.get("parity", {}).get("max_abs_diff")) for k, v in (d.get("variants") or {}).items()})
except Exception as exc:
    print("no json:", exc)
PY
  grep -v "Warn\|sparse_csr\|OMP_NUM\|\*\*\*\*\|NCCL version" gpurun_out/bench_${name}_n$n.err | tail -4 | cut -c1-300
}
BARGS="--no-variants --no-e2e" run calib 8 TCGNN_EXCHANGE=auto
BARGS="--no-variants --no-e2e" run calib_x1 8 TCGNN_EXCHANGE=auto TCGNN_CALIBRATE_PRODUCTS=1

