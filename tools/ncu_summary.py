"""Extract the judged subset of an `ncu -i rep --page raw --csv` export into a small `metric,value,unit` csv (profiles/r1_*_metrics.csv).
usage: ncu_summary.py raw.csv out.csv"""
import csv
import sys

KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpc__cycles_elapsed.max", "gpu__time_duration.sum", "l1tex__t_sector_hit_rate.pct",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__stack_size", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum")
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, vals = rows[0], rows[1], rows[2]
with open(sys.argv[2], "w") as f:
    f.write("metric,value,unit\n")
    for h, u, v in sorted(zip(hdr, units, vals)):
        if h in KEEP or (h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")):
            f.write(f"{h},{v},{u}\n")
