#!/usr/bin/env bash
# L2 evict_last hint on the feature-row gathers: off (default build) vs on (variants/xhint), every operator and size.
ITEMS="spmm:reddit-like-rmat wspmm:reddit-like-rmat sddmm:reddit-like-rmat agnn:reddit-like-rmat spmm:reddit-like-uniform sddmm:reddit-like-uniform spmm:products-like-rmat sddmm:products-like-rmat agnn:products-like-rmat spmm:rmat-10m-200m"
timeout 300 python tools/ab.py --tag nohint $ITEMS 2>&1 | grep "min_ms\|rror"
LD_LIBRARY_PATH=$PWD/variants/xhint timeout 300 python tools/ab.py --tag hint $ITEMS 2>&1 | grep "min_ms\|rror"
