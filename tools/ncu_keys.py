#!/usr/bin/env python
"""Print the handful of ncu metrics that decide what to tune next from a `--set full` report.
Usage: python tools/ncu_keys.py report.ncu-rep [more substrings...]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum ",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread ",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum ", "dram__bytes_write.sum ", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "launch__occupancy_limit",
        "lts__t_bytes.sum ", "sm__inst_executed_pipe_fp64"]


def main():
    rep = sys.argv[1]
    keys = KEYS + sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("==", vals[hdr.index("Kernel Name")][:100])
        for h, u, v in zip(hdr, units, vals):
            if any((h + " ").startswith(k) or k in h + " " for k in keys):
                if v in ("0", "0.000000", "no data"):
                    continue
                print(f"   {h:88s} {u:14s} {v}")


if __name__ == "__main__":
    main()
