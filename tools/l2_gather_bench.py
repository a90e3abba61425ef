#!/usr/bin/env python3
"""Builds (if needed) and runs tools/l2_gather_bench.cu on the GPU box, keeps every configuration's line and derives
the L2 -> SM gather ceiling bench.py reports `roofline.l2_gather` against.

    python tools/l2_gather_bench.py [--quick] [--out gpurun_out/l2_gather]

Writes <out>.jsonl (all lines) and <out>_peak.json = {"bytes_per_clk": best ldgsts figure on an L2-resident working
set, "gbs": ..., "at": that line, "reddit_sized": best ldgsts line at 119 MB, "gather4": best gather4 lines}.  Copy
the latter to profiles/l2_gather_peak.json (bench.py reads it from there)."""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tools", "l2_gather_bench.cu")
BIN = os.path.join(ROOT, "tools", "build", "l2_gather_bench")


def build():
    if os.path.exists(BIN) and os.path.getmtime(BIN) >= os.path.getmtime(SRC):
        return
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
                           "-o", BIN, SRC])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "l2_gather"))
    ap.add_argument("--from-jsonl", default=None, help="re-derive <out>_peak.json from a kept sweep (no GPU needed)")
    args = ap.parse_args()
    if args.from_jsonl:
        with open(args.from_jsonl) as fh:
            lines = [json.loads(ln) for ln in fh if ln.startswith("{")]
        rc = 0
    else:
        build()
        res = subprocess.run([BIN] + (["--quick"] if args.quick else []), stdout=subprocess.PIPE, text=True, timeout=1500)
        rc = res.returncode
        lines = [json.loads(ln) for ln in res.stdout.splitlines() if ln.startswith("{")]
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out + ".jsonl", "w") as fh:
            for ln in lines:
                fh.write(json.dumps(ln) + "\n")
    runs = [ln for ln in lines if "mode" in ln]
    if rc != 0 or not runs:
        print(f"l2_gather_bench failed (rc={rc}); {len(runs)} lines kept", file=sys.stderr)
        return 1

    def best(mode, lo, hi):
        c = [r for r in runs if r["mode"] == mode and lo <= r["working_set_mb"] <= hi]
        return max(c, key=lambda r: r["bytes_per_clk"]) if c else None

    resident = best("ldgsts", 0, 64)
    peak = {"bytes_per_clk": resident["bytes_per_clk"], "gbs": resident["gbs"], "at": resident,
            "what": "best of the ldgsts sweep on an L2-resident working set (<= 64 MB), random 512-byte rows, whole chip",
            "reddit_sized": best("ldgsts", 110, 130), "beyond_l2": best("ldgsts", 400, 1e9),
            "gather4": {"resident": best("gather4", 0, 64), "reddit_sized": best("gather4", 110, 130)},
            "gather4_wide": {"resident": best("gather4_wide", 0, 64), "reddit_sized": best("gather4_wide", 110, 130)},
            "checks": [ln for ln in lines if "check" in ln],
            # best ldgsts figure per working-set size: the ceiling for a feature matrix of that size
            "curve": [{"working_set_mb": ws, "bytes_per_clk": best("ldgsts", ws - 0.01, ws + 0.01)["bytes_per_clk"],
                       "gbs": best("ldgsts", ws - 0.01, ws + 0.01)["gbs"]}
                      for ws in sorted({r["working_set_mb"] for r in runs if r["mode"] == "ldgsts"})]}
    with open(args.out + "_peak.json", "w") as fh:
        json.dump(peak, fh, indent=1)
    print(json.dumps({k: peak[k] for k in ("bytes_per_clk", "gbs")}))
    return 0


if __name__ == "__main__":
    sys.exit(main())
