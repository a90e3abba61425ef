#!/usr/bin/env bash
mkdir -p gpurun_out
echo "=== diag products"; timeout 600 python tools/diag_sddmm.py products-like-rmat 256 3 2>&1 | grep -v Warn | tail -60 | tee gpurun_out/diag_products.txt
echo "=== diag products team1"; TCGNN_SPMM_TEAM=1 timeout 600 python tools/diag_sddmm.py products-like-rmat 256 3 2>&1 | grep -v Warn | grep -E "weighted|TOTAL" | tee gpurun_out/diag_products_team1.txt
echo "=== diag c5"; timeout 900 python tools/diag_sddmm.py rmat-10m-200m 256 3 2>&1 | grep -v Warn | tail -70 | tee gpurun_out/diag_c5.txt
echo "=== diag mid D=256"; timeout 600 python tools/diag_sddmm.py rmat:2000000:40000000 256 4 2>&1 | grep -v Warn | tail -40 | tee gpurun_out/diag_mid256.txt
echo "=== diag mid D=128"; timeout 600 python tools/diag_sddmm.py rmat:2000000:40000000 128 4 2>&1 | grep -v Warn | tail -40 | tee gpurun_out/diag_mid128.txt
echo "=== racecheck small"; timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python tools/diag_sddmm.py rmat:30000:600000 256 1 2>&1 | grep -v Warn | tail -40 | tee gpurun_out/racecheck.txt
