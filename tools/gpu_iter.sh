#!/usr/bin/env bash
# iteration: spmm parity + quick timings
timeout 600 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_vs_reference.py -m gpu -q --timeout 300 -x 2>&1 | tail -5
for wl in reddit-like-uniform reddit-like-rmat; do
for ab in 0 7 1 2; do
  TCGNN_ABLATE=$ab timeout 300 python tools/quick.py --workload $wl --iters 3 2>&1 | tail -1
done; done
timeout 300 python tools/quick.py --workload products-like-rmat --iters 3 2>&1 | tail -1
timeout 300 python tools/quick.py --workload citeseer-like --iters 20 2>&1 | tail -1
