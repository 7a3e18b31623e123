"""Condense an .ncu-rep (ncu --set full, one kernel launch) into the text summary kept under profiles/.
usage: ncu_summary.py <report.ncu-rep> [title]"""
import csv, subprocess, sys
rep = sys.argv[1]
title = sys.argv[2] if len(sys.argv) > 2 else rep
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units, v = rows[0], rows[1], rows[2]
m = {n: (v[i], units[i]) for i, n in enumerate(h)}
keys = [
    ("Kernel Name", "kernel"), ("Block Size", "block"), ("Grid Size", "grid"),
    ("gpu__time_duration.sum", "duration"), ("launch__registers_per_thread", "registers/thread"),
    ("launch__occupancy_limit_registers", "occupancy limit (registers) [CTAs/SM]"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy [% of 64 warps/SM]"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active [% of peak]"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 instructions issued [% of peak]"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy [%]"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "warp execution efficiency [active threads / 32-wide instruction]"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("dram__bytes_read.sum", "DRAM bytes read"), ("dram__bytes_write.sum", "DRAM bytes written"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput [% of peak]"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate [%]"), ("l1tex__t_sector_hit_rate.pct", "L1 hit rate [%]"),
    ("sass__inst_executed_local_loads", "local (spill) loads"), ("sass__inst_executed_local_stores", "local (spill) stores"),
]
print("# " + title)
for k, label in keys:
    if k in m:
        print("%-72s %s %s" % (label, m[k][0], m[k][1]))
print("\nwarp stall reasons (average warps stalled per issue-active cycle):")
st = [(float(m[k][0]), k) for k in m if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
for val, k in sorted(st, reverse=True)[:9]:
    print("  %-28s %.3f" % (k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], val))
