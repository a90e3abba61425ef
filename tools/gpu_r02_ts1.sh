#!/usr/bin/env bash
# Register gathers + A operand in tensor memory: layout probe, MMA issue rate, SpMM parity, timings.
mkdir -p gpurun_out
echo "=== probe"; timeout 120 python -m pytest tests/test_gpu_umma_layouts.py -x -q -k tensor_memory 2>&1 | tail -5
cat gpurun_out/umma_probe.txt 2>/dev/null | tail -3
echo "=== issue rate"; timeout 120 python tools/umma_bench.py ts 2>&1 | tail -8
echo "=== parity (TS)"; TCGNN_SPMM_TS=1 timeout 300 python -m pytest tests/test_gpu_spmm.py -x -q 2>&1 | tail -8
ITEMS="spmm:reddit-like-rmat wspmm:reddit-like-rmat spmm:reddit-like-uniform spmm:products-like-rmat"
echo "=== timings"
timeout 200 python tools/ab.py --tag base $ITEMS 2>&1 | grep min_ms
TCGNN_SPMM_TS=1 TCGNN_SPMM_TS_GROUPS=3 timeout 200 python tools/ab.py --tag ts3 $ITEMS 2>&1 | grep "min_ms\|rror"
TCGNN_SPMM_TS=1 TCGNN_SPMM_TS_GROUPS=4 timeout 200 python tools/ab.py --tag ts4 $ITEMS 2>&1 | grep "min_ms\|rror"
