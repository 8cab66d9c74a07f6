"""Summarises an .ncu-rep (ncu --set full) as a markdown table of the metrics the design discussion uses.
Usage: python scripts/ncu_summary.py report.ncu-rep "<title line>" > profiles/<name>.md"""
import csv, subprocess, sys
rep, title = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h, units, rows = r[0], r[1], r[2:]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_bytes.sum.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"]
ki = h.index("Kernel Name")
names = [row[ki].split("(")[0].replace("gp::", "") for row in rows]
print(title)
print()
print("| metric | unit | " + " | ".join(names) + " |")
print("|---|---|" + "---|" * len(names))
for m in want:
    if m not in h: continue
    i = h.index(m)
    print(f"| {m} | {units[i]} | " + " | ".join(row[i] for row in rows) + " |")
