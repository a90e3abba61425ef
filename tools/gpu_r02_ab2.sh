#!/usr/bin/env bash
# Does the LDGSTS flavour (zero-fill size operand, L2 cache hint) cost gather throughput?  Timing only: the plain
# variants read row 0 for padding rows, so their results are wrong on padded tiles.
ITEMS="spmm:reddit-like-rmat spmm:reddit-like-uniform"
timeout 200 python tools/ab.py --tag base $ITEMS 2>&1 | grep min_ms
for v in plain1 plain2; do
  LD_LIBRARY_PATH=$PWD/variants/$v timeout 200 python tools/ab.py --tag $v $ITEMS 2>&1 | grep min_ms
  LD_LIBRARY_PATH=$PWD/variants/$v TCGNN_SPMM_TEAM=2 timeout 200 python tools/ab.py --tag $v-team2 $ITEMS 2>&1 | grep min_ms
done
