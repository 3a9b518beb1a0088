#!/usr/bin/env python
"""ncu_summary.py REPORT.ncu-rep [metric ...] -- one line per captured launch with the counters the profiles/ notes quote"""
import csv
import subprocess
import sys

WANT = ["Kernel Name", "Grid Size", "Block Size", "Registers Per Thread", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__sectors_read.sum", "lts__t_requests_srcunit_tex.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
        "lts__t_sectors_srcunit_tex_op_write.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct"]


def main():
    rep = sys.argv[1]
    want = WANT + sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("---")
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:80s} {r[i]} {units[i]}")


if __name__ == "__main__":
    main()
