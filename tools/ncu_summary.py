"""Compact per-launch summary (CSV) of an ncu --set full report: python tools/ncu_summary.py rep.ncu-rep > out.csv"""
import csv
import subprocess
import sys

METRICS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_sector_hit_rate.pct",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
# a .ncu-rep report, or the `ncu -i rep --page raw --csv` dump of one (what a GPU box sends back when the report is too large)
if sys.argv[1].endswith(".csv"):
    raw = open(sys.argv[1]).read()
else:
    raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
w = csv.writer(sys.stdout)
w.writerow(["id", "kernel", "grid", "block"] + [f"{m} [{units[idx[m]]}]" for m in METRICS if m in idx])
only_team = len(sys.argv) > 2 and sys.argv[2] == "--team-only"
for r in data:
    if only_team and "team::" not in r[idx["Kernel Name"]]:
        continue
    name = r[idx["Kernel Name"]].split("(")[0].replace("team::", "").replace("void ", "")
    w.writerow([r[idx["ID"]], name, r[idx["Grid Size"]], r[idx["Block Size"]]] + [r[idx[m]] for m in METRICS if m in idx])
