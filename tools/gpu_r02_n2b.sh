#!/usr/bin/env bash
mkdir -p gpurun_out
echo "=== nccl parity (2 ranks)"
timeout 900 python -m pytest tests/test_gpu_sharding.py -m gpu -q --timeout 600 -k two_ranks 2>&1 | grep -v Warn | tail -25 | tee gpurun_out/test_nccl_n2.log
for mode in 1 0; do
echo "=== bench overlap graph=$mode (N=2)"
TCGNN_EXCHANGE_GRAPH=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
   bench.py --gpus 2 --steps 20 --warmup 5 --no-variants --no-e2e > gpurun_out/bench_overlap_g${mode}_n2.json 2> gpurun_out/bench_overlap_g${mode}_n2.err
echo rc=$?; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_overlap_g${mode}_n2.json").read())
    print({k: d.get(k) for k in ("ms_per_step", "min_ms", "parity")})
except Exception as exc:
    print("no json:", exc)
PY
grep -v "Warn\|sparse_csr\|OMP_NUM\|\*\*\*\*" gpurun_out/bench_overlap_g${mode}_n2.err | tail -8 | cut -c1-300
done
