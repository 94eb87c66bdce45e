"""Warp-instructions executed per source line of an .ncu-rep (needs -lineinfo + --import-source on).
usage: python scripts/ncu_instr.py gpurun_out/x.ncu-rep [top_n]"""
import csv, subprocess, io, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None; fname = ''; data = []
for r in rows:
    if len(r) == 2 and r[0] == "File Path": fname = r[1].split('/')[-1]; continue
    if len(r) > 6 and r[0] == "Line No": hdr = r; continue
    if hdr and len(r) == len(hdr) and r[2] == '-':
        try: ie = int(r[hdr.index("Instructions Executed")]); te = int(r[hdr.index("Thread Instructions Executed")])
        except Exception: continue
        data.append((ie, te, fname, r[0], r[1].strip()[:100]))
tot = sum(d[0] for d in data)
print('total warp-instructions', tot, ' avg active threads %.1f' % (sum(d[1] for d in data) / max(tot, 1)))
for ie, te, f, ln, src in sorted(data, key=lambda d: -d[0])[:top]:
    print("%10d %5.1f%% act %4.1f %s:%s %s" % (ie, 100 * ie / tot, te / max(ie, 1), f[:14], ln, src))
