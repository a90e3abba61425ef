#!/usr/bin/env bash
# multi-GPU check: sharding tests (incl. the world_size>1 NCCL test) and the strong-scaling bench line
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
true
for wl in ${WORKLOADS:-reddit-like-rmat}; do
for n in ${NS:-$N}; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 --workload $wl --no-cpu-baseline > gpurun_out/bench_${wl}_n$n.json 2> gpurun_out/bench_${wl}_n$n.err
echo "torchrun exit code $?"
python - <<PY
import json
try:
    d = [json.loads(l) for l in open("gpurun_out/bench_${wl}_n$n.json") if l.startswith("{")][-1]
    print("$wl", "N=$n", "ms/step", d["ms_per_step"], "Gedges/s", round(d["value"] / 1e9, 2), "kernel_ms", d["roofline"]["kernel_ms"], "e2e_ms", d["e2e"]["ms_per_step"])
except Exception as e:
    print("$wl N=$n FAILED", e)
PY
echo "exit code $? (of the summary)"; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/bench_${wl}_n$n.err | tail -15 | cut -c1-400
done; done
