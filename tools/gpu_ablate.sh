#!/usr/bin/env bash
wl=reddit-like-uniform
echo "--- mma trace noinc (8+32)";   TCGNN_ABLATE=40  TCGNN_TRACE=2000 timeout 300 python tools/quick.py --workload $wl --iters 1 2>&1 | grep -E "trace|min_ms" | head -20
echo "--- mma trace all ablated (15+32)"; TCGNN_ABLATE=47 TCGNN_TRACE=2000 timeout 300 python tools/quick.py --workload $wl --iters 1 2>&1 | grep -E "trace|min_ms" | head -20
