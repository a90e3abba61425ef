#!/usr/bin/env bash
# The default line under torchrun with the C5 variant forced on (normally N = 1 and N = 8 only): measured-time partition
# + overlapped exchange on the 10 M / 200 M graph, the e2e section at N > 1 and the guarded optional sections.
mkdir -p gpurun_out
N=${1:-2}
TCGNN_BENCH_VARIANTS=all timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --gpus $N > gpurun_out/bench_default_allvariants_n$N.json 2> gpurun_out/bench_default_allvariants_n$N.err
echo "rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_default_allvariants_n$N.json").read())
print({k: d.get(k) for k in ("ms_per_step", "min_ms", "incomplete")}, "bit_exact", d["parity"]["bit_exact"], "e2e", (d.get("e2e") or {}).get("ms_per_step"))
print("partition", d.get("partition"))
for k, v in (d.get("variants") or {}).items():
    print(k, {a: v.get(a) for a in ("ms_per_step", "parity", "partition", "failed")})
PY
grep -v "Warn\|sparse_csr\|OMP_NUM\|\*\*\*\*\|NCCL version" gpurun_out/bench_default_allvariants_n$N.err | tail -6 | cut -c1-300
