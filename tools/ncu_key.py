#!/usr/bin/env python3
"""Key metrics of one ncu report: usage  ncu -i rep --page raw --csv | python tools/ncu_key.py"""
import csv
import sys

rows = list(csv.reader(sys.stdin))
h, u, v = rows[0], rows[1], rows[-1]
d = dict(zip(h, v))
un = dict(zip(h, u))
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum",
        "smsp__inst_executed_op_tma_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum"]
for k in want:
    if k in d:
        print(f"{k:75s} {d[k]:>18s} {un[k]}")
st = [k for k in h if "smsp__average_warps_issue_stalled" in k and k.endswith("_per_issue_active.ratio")]
def f(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0
print("stall reasons (warps per issue-active cycle):")
for k in sorted(st, key=lambda k: -f(d[k]))[:9]:
    print("   %-40s %s" % (k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), d[k]))
