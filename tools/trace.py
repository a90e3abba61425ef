#!/usr/bin/env python3
"""Summarise a TCGNN_TRACE dump (block 0 timeline of spmm_tc_kernel): python tools/trace.py file [first last]"""
import sys
import numpy as np

raw = np.fromfile(sys.argv[1], dtype=np.int64)
t = raw[:3 * 512 * 8].reshape(3, 512, 8)
cta = raw[3 * 512 * 8:3 * 512 * 8 + 160 * 4].reshape(-1, 4)
epi = raw[3 * 512 * 8 + 160 * 4:].reshape(512, 8)
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 100
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 400
names = {
    0: ["full ready", "info read", "MMAs issued", "commit issued"],
    1: ["top", "meta ready", "B values+flags", "prev landed", "prev published", "slot free", "gathers issued", "B stored"],
    2: ["top", "slot free", "TMA issued"],
}
import os
if os.environ.get("TCGNN_SPMM_TS", "0") != "0":   # register-gather pipeline: different points
    names[1] = ["top", "meta ready", "B values+flags", "loads issued", "records released", "slot free", "tcgen05.st issued", "published"]
for role, label in ((0, "MMA warp (per stage)"), (1, "producer warp 0 (per own stage)"), (2, "meta loader (per stage)")):
    r = t[role, lo:hi]
    n = len(names[role])
    if role == 1:
        r = t[role, 10:85]                      # own stages only: far fewer samples per CTA
        keep = (r > 0).all(axis=0)             # points this pipeline shape never records (lag 0: "prev landed")
        r = r[:, keep]
        names[1] = [nm for nm, k in zip(names[1], keep) if k]
        n = len(names[1])
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

cta = cta[cta[:, 0] > 0]
if len(cta):
    cyc, tiles, wins = cta[:, 0].astype(float), cta[:, 1].astype(float), cta[:, 2].astype(float)
    print(f"per-CTA: {len(cta)} CTAs, cycles min {cyc.min():.0f} mean {cyc.mean():.0f} max {cyc.max():.0f}; "
          f"tiles {tiles.min():.0f}..{tiles.max():.0f}; windows {wins.min():.0f}..{wins.max():.0f}")
    for i in np.argsort(-cyc)[:5]:
        print(f"   CTA {i:3d}: {cyc[i]:12.0f} cycles  {tiles[i]:8.0f} tiles  {wins[i]:8.0f} windows  {cyc[i] / tiles[i]:7.1f} cyc/tile  {cyc[i] / max(wins[i], 1):8.1f} cyc/window")
    A = np.stack([tiles, wins], 1)
    coef, *_ = np.linalg.lstsq(A, cyc, rcond=None)
    print(f"   least squares: cycles ~ {coef[0]:.1f} * tiles + {coef[1]:.1f} * windows")

e = epi[lo:hi]
e = e[(e[:, [0, 1, 2, 4, 5]] > 0).all(axis=1)]
if len(e) > 2:
    print(f"epilogue warp 0 (per window): {len(e)} samples, period {np.diff(e[:, 0]).mean():.0f} cycles")
    last = np.where(e[:, 3] > 0, e[:, 3], e[:, 2])
    for name, d in (("wait acc_full", e[:, 1] - e[:, 0]), ("tcgen05.ld (m=0)", e[:, 2] - e[:, 1]),
                    ("stores (+ld m=1)", e[:, 4] - e[:, 2]), ("fence+arrive", e[:, 5] - e[:, 4])):
        print(f"   {name:>18s} mean {d.mean():7.0f}  p50 {np.median(d):7.0f}  p90 {np.percentile(d, 90):7.0f}")
