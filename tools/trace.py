#!/usr/bin/env python3
"""Summarise a TCGNN_TRACE dump (block 0 timeline of spmm_tc_kernel): python tools/trace.py file [first last]"""
import sys
import numpy as np

t = np.fromfile(sys.argv[1], dtype=np.int64).reshape(3, 512, 8)
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 100
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 400
names = {
    0: ["full ready", "info read", "MMAs issued", "commit issued"],
    1: ["top", "meta ready", "B values+flags", "prev landed", "prev published", "slot free", "gathers issued", "B stored"],
    2: ["top", "slot free", "TMA issued"],
}
for role, label in ((0, "MMA warp (per stage)"), (1, "producer warp 0 (per own stage)"), (2, "meta loader (per stage)")):
    r = t[role, lo:hi]
    n = len(names[role])
    ok = (r[:, :n] > 0).all(axis=1)
    r = r[ok]
    if len(r) < 2:
        print(label, ": no data")
        continue
    period = np.diff(r[:, 0]).mean()
    print(f"{label}: {len(r)} samples, period {period:.0f} cycles")
    for i in range(1, n):
        d = r[:, i] - r[:, i - 1]
        print(f"   {names[role][i - 1]:>16s} -> {names[role][i]:<16s} mean {d.mean():7.0f}  p50 {np.median(d):7.0f}  p90 {np.percentile(d, 90):7.0f}")
    d = r[1:, 0] - r[:-1, n - 1]
    print(f"   {names[role][n - 1]:>16s} -> next {names[role][0]:<11s} mean {d.mean():7.0f}  p50 {np.median(d):7.0f}  p90 {np.percentile(d, 90):7.0f}")
