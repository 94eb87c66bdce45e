"""Key metrics of every kernel in an .ncu-rep.  usage: python scripts/ncu_metrics.py rep.ncu-rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h, u = rows[0], rows[1]
K = ["gpu__time_duration.sum", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
     "smsp__inst_executed.sum", "launch__waves_per_multiprocessor", "launch__registers_per_thread", "launch__grid_size", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
for v in rows[2:]:
    print("==", v[h.index("Kernel Name")][:60])
    for k in K:
        if k in h: print("   %-62s %s %s" % (k, v[h.index(k)], u[h.index(k)]))
