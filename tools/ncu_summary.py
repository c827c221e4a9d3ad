#!/usr/bin/env python
"""Selected rows of an `ncu -i <file>.ncu-rep --page raw --csv` dump, one block per captured launch (evidence tool).

usage: ncu_summary.py raw.csv ["comment line" ...]  > profiles/<name>.txt
The .ncu-rep files are megabytes of scratch (gpurun_out/); the rows the design discussion cites are what gets committed.
"""
import csv
import re
import sys

KEEP = [
    "Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_static",
    "sm__warps_active.avg.per_cycle_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed_per_inst_executed.pct",
    "sm__cycles_elapsed.max", "smsp__cycles_active.avg",
]
PATTERNS = [r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio", r"smsp__average_warp.*latency.*"]


def main():
    rows = list(csv.reader(open(sys.argv[1], newline="")))
    for c in sys.argv[2:]:
        print("# " + c)
    head = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units = rows[head], rows[head + 1]
    for r in rows[head + 2:]:
        if len(r) != len(names):
            continue
        d = dict(zip(names, zip(units, r)))
        keys = [k for k in KEEP if k in d] + sorted(k for k in d if any(re.fullmatch(p, k) for p in PATTERNS))
        for k in keys:
            u, v = d[k]
            print(f"{k:<100s} {u:<16s} {v}")
        print()


if __name__ == "__main__":
    main()
