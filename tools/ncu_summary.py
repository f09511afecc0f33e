"""Selected counters per kernel from an ncu report: `ncu -i X.ncu-rep --page raw --csv > raw.csv; python tools/ncu_summary.py raw.csv`.
One line per (kernel, metric): report,kernel,metric,value,unit — the format of profiles/*_ncu_full_summary.csv."""
import csv
import re
import sys

WANT = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
name = hdr.index("Kernel Name")
report = sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]
print("report,kernel,metric,value,unit")
seen = set()
for r in rows[2:]:
    k = re.sub(r"\(.*", "", r[name]).split("::")[-1].replace("void ", "")
    if k in seen:          # first captured launch of each kernel
        continue
    seen.add(k)
    for m in WANT:
        if m in hdr:
            print(f"{report},{k},{m},{r[hdr.index(m)]},{units[hdr.index(m)]}")
