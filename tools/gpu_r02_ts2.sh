#!/usr/bin/env bash
# Register gathers + A operand in tensor memory: SpMM parity, then timings.
mkdir -p gpurun_out
echo "=== parity (TS)"; TCGNN_SPMM_TS=1 timeout 300 python -m pytest tests/test_gpu_spmm.py -x -q 2>&1 | tail -8
ITEMS="spmm:reddit-like-rmat wspmm:reddit-like-rmat spmm:reddit-like-uniform spmm:products-like-rmat"
echo "=== timings"
TCGNN_SPMM_TS=1 TCGNN_SPMM_TS_GROUPS=3 timeout 200 python tools/ab.py --tag ts3 $ITEMS 2>&1 | grep "min_ms\|rror"
TCGNN_SPMM_TS=1 TCGNN_SPMM_TS_GROUPS=4 timeout 200 python tools/ab.py --tag ts4 $ITEMS 2>&1 | grep "min_ms\|rror"
