"""Summaries for profiles/: (1) launch list shares from an `ncu --metrics gpu__time_duration.sum` CSV,
(2) key metrics of `ncu --set full` reports.  usage: python scripts/ncu_summary.py launches.csv rep1.ncu-rep ..."""
import collections, csv, io, json, subprocess, sys

def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    h = rows[0]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        agg.setdefault(name, []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    out = []
    for n, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        out.append({"kernel": n, "launches": len(v), "avg_us": round(sum(v) / len(v) / 1e3, 2), "share": round(sum(v) / tot, 4)})
    return {"total_ms": round(tot / 1e6, 3), "kernels": out}

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "launch__waves_per_multiprocessor", "smsp__cycles_active.avg"]

def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, units, v = rows[0], rows[1], rows[-1]
    d = {"kernel": v[h.index("Kernel Name")].split("(")[0]}
    for i, n in enumerate(h):
        if n in WANT:
            d[n] = "%s %s" % (v[i], units[i])
    return d

if __name__ == "__main__":
    res = {"launch_list": launches(sys.argv[1]), "full": [full(p) for p in sys.argv[2:]]}
    print(json.dumps(res, indent=1))
