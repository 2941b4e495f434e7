#!/usr/bin/env python
"""Per-source-line instruction / stall-sample breakdown of one kernel in an .ncu-rep (compiled with -lineinfo):
   python tools/ncu_lines.py report.ncu-rep file.cu [top]"""
import collections, csv, subprocess, sys
rep, fname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
by = collections.OrderedDict(); tot = ts = 0; cur = None; hdr = None; infile = False
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        infile = r[1].endswith(fname); hdr = None; continue
    if "Line No" in r:
        hdr = r; iL = r.index("Line No"); iE = r.index("Instructions Executed"); iS = r.index("# Samples"); iSrc = r.index("Source"); continue
    if hdr is None or len(r) < len(hdr): continue
    if r[iL] != "":
        cur = (int(r[iL]), r[iSrc].strip()[:74]) if infile else ("other", "")
    try: n = int(r[iE]); smp = int(r[iS])
    except ValueError: continue
    d = by.setdefault(cur, [0, 0]); d[0] += n; d[1] += smp; tot += n; ts += smp
print("total warp instructions", tot, "samples", ts)
for k, (n, smp) in sorted(by.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"smp {100*smp/ts:5.1f}%  inst {100*n/tot:5.1f}%  L{k[0]}: {k[1]}")
