#!/usr/bin/env python
"""Per-source-line instruction / stall-sample breakdown of one kernel in an .ncu-rep (compiled with -lineinfo):
   python tools/ncu_lines.py report.ncu-rep [systems_per_launch] [min_pct]"""
import collections, csv, subprocess, sys
rep = sys.argv[1]
nsys = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
minp = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
by = collections.OrderedDict(); tot = ts = 0; hdr = None; fn = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fn = r[1].split("/")[-1]; hdr = None; continue
    if "Line No" in r:
        hdr = r; iL = r.index("Line No"); iE = r.index("Instructions Executed"); iS = r.index("# Samples"); iSrc = r.index("Source"); continue
    if hdr is None or len(r) < len(hdr) or r[iL] == "": continue   # CUDA-level rows only (SASS rows repeat them)
    try: n = int(r[iE]); smp = int(r[iS])
    except ValueError: continue
    by[(fn, int(r[iL]), r[iSrc].strip()[:86])] = (n, smp); tot += n; ts += smp
print(f"warp instructions {tot} ({tot / nsys:.0f} per unit), samples {ts}")
for k, (n, smp) in sorted(by.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if n > tot * minp / 100 or smp > ts * minp / 100:
        print(f"{k[0][:16]:16s} L{k[1]:4d} inst {n / nsys:8.0f} {100 * n / tot:5.1f}%  smp {100 * smp / ts:5.1f}%  {k[2]}")
