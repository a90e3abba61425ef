#!/usr/bin/env python3
"""Summarise an ncu report (run here, no GPU needed): python tools/ncu_summary.py rep.ncu-rep > profiles/x.txt"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.avg ", "sm__cycles_elapsed.avg.per_second",
        "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum ", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum ", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum ",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed.sum.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor", "smsp__inst_executed_op_ldgsts", "sm__inst_executed_pipe_uniform"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
        print(f"== kernel: {name}")
        stalls = []
        for i, h in enumerate(hdr):
            if any(h.startswith(k) or (k.endswith(" ") and h == k.strip()) for k in KEYS):
                print(f"{h:90s} {vals[i]:>18s} {units[i]}")
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
                try:
                    stalls.append((float(vals[i]), h))
                except ValueError:
                    pass
        print("-- warp stall reasons (warps per issue-active cycle), top 8")
        for v, h in sorted(stalls, reverse=True)[:8]:
            print(f"{v:8.3f}  {h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')}")


if __name__ == "__main__":
    main()
