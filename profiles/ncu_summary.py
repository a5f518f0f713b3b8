#!/usr/bin/env python
"""Summarise an .ncu-rep: per kernel launch the metrics the roofline discussion needs.
usage: python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [> profiles/rNN_xxx.txt]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
STALLS = ["barrier", "branch_resolving", "dispatch_stall", "drain", "lg_throttle", "long_scoreboard", "math_pipe_throttle", "membar",
          "mio_throttle", "misc", "no_instruction", "not_selected", "selected", "short_scoreboard", "sleeping", "tex_throttle", "wait"]
KEYS += [f"smsp__average_warps_issue_stalled_{x}_per_issue_active.ratio" for x in STALLS]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    for r in data:
        print(f"== [{r[col['ID']]}] {r[col['Kernel Name']][:90]}  grid={r[col['Grid Size']]} block={r[col['Block Size']]}")
        for k in KEYS:
            if k in col:
                print(f"   {k:75s} {r[col[k]]:>16s} {units[col[k]]}")


if __name__ == "__main__":
    main(sys.argv[1])
