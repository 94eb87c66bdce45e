"""Per-source-line warp-stall sample summary of an .ncu-rep (needs -lineinfo + --import-source on).
usage: python scripts/ncu_lines.py gpurun_out/x.ncu-rep [top_n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
data = []; fname = ""
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 6 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":   # source-line summary rows (SASS rows carry an address)
        try: s = int(r[hdr.index("# Samples")])
        except ValueError: continue
        stalls = {hdr[i].replace("stall_", ""): int(r[i]) for i in range(len(hdr)) if hdr[i].startswith("stall_") and "Not Issued" not in hdr[i] and r[i].isdigit() and int(r[i]) > 0}
        data.append((s, fname, r[0], r[1].strip()[:100], stalls))
tot = sum(d[0] for d in data)
print("total samples", tot)
for s, f, ln, src, st in sorted(data, key=lambda d: -d[0])[:top]:
    top3 = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print("%6d %5.1f%% %s:%-4s %-28s %s" % (s, 100.0 * s / max(tot, 1), f[:14], ln, " ".join("%s=%d" % kv for kv in top3), src))
