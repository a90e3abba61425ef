mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sharding.py -m gpu -q --timeout 300 -x 2>&1 | tail -15
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 tools/exchange_probe.py 2>&1 | grep "^\[" | tee gpurun_out/exchange_probe_n4.txt
for mode in p2p p2p2; do
echo "--- bench N=4 $mode"
TCGNN_EXCHANGE=$mode timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/m4_$mode.err | grep "^{" > gpurun_out/bench_n4_$mode.json
python -c "
import json; d=json.load(open('gpurun_out/bench_n4_$mode.json')); print('$mode ms/step', d['ms_per_step'], 'Gedges/s', round(d['value']/1e9,1), 'kernel_ms', d['roofline']['kernel_ms'])"
grep -v "^\*\*\*\|OMP_NUM" gpurun_out/m4_$mode.err | tail -3 | cut -c1-300
done
