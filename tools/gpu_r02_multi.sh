#!/usr/bin/env bash
# Multi-GPU pass (gpurun --gpus N): NCCL parity test of every exchange mode, then bench lines with the overlapped
# exchange and with the round-1 gather-then-compute exchange.  usage: gpu_r02_multi.sh N [extra]
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -9
echo "=== nccl parity (2 ranks)"
timeout 900 python -m pytest tests/test_gpu_sharding.py -m gpu -q --timeout 600 -k two_ranks 2>&1 | grep -v Warn | tail -15 | tee gpurun_out/test_nccl_n$N.log
run() {  # name, env..., then bench args
  name=$1; shift
  echo "=== bench $name (N=$N)"
  env "$@" timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $N --steps 20 --warmup 5 $BARGS > gpurun_out/bench_${name}_n$N.json 2> gpurun_out/bench_${name}_n$N.err
  echo "rc=$?"; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${name}_n$N.json"))
    print({k: d.get(k) for k in ("ms_per_step", "value", "parity", "exchange")})
    print("e2e", (d.get("e2e") or {}).get("ms_per_step"), "variants", {k: (v.get("ms_per_step"), v.get("parity", {}).get("max_abs_diff"), v.get("exchange")) for k, v in (d.get("variants") or {}).items()})
except Exception as exc:
    print("no json:", exc)
PY
  tail -4 gpurun_out/bench_${name}_n$N.err | cut -c1-300
}
BARGS="--no-variants" run overlap TCGNN_EXCHANGE=auto
BARGS="--no-variants --no-e2e" run gather_p2p TCGNN_EXCHANGE=p2p
BARGS="--no-variants --no-e2e" run gather_p2p2 TCGNN_EXCHANGE=p2p2
if [ "$2" = "c5" ]; then
  BARGS="--workload rmat-10m-200m --steps 10 --warmup 3 --no-e2e" run c5_overlap TCGNN_EXCHANGE=auto
  BARGS="--workload rmat-10m-200m --steps 10 --warmup 3 --no-e2e" run c5_gather TCGNN_EXCHANGE=p2p2
fi
