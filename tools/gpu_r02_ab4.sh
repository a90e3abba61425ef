#!/usr/bin/env bash
# After trimming the row-address arithmetic of the gathers (one IMAD.WIDE per row): parity of the touched kernels, timings.
timeout 400 python -m pytest tests/test_gpu_spmm.py tests/test_gpu_sddmm.py tests/test_gpu_fused_ops.py -x -q 2>&1 | tail -4
ITEMS="spmm:reddit-like-rmat wspmm:reddit-like-rmat sddmm:reddit-like-rmat agnn:reddit-like-rmat spmm:reddit-like-uniform sddmm:reddit-like-uniform spmm:products-like-rmat sddmm:products-like-rmat agnn:products-like-rmat spmm:rmat-10m-200m"
timeout 300 python tools/ab.py --tag trimmed $ITEMS 2>&1 | grep "min_ms\|rror"
