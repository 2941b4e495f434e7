#!/usr/bin/env python
"""Summarise ncu outputs into profiles/: (1) a launch list CSV (--metrics gpu__time_duration.sum) -> per-kernel count /
total / share table; (2) .ncu-rep full captures -> the handful of metrics DESIGN.md and bench.py cite.

    python tools/ncu_summary.py launches gpurun_out/r01_launches.csv profiles/r01_launches_summary.md
    python tools/ncu_summary.py full gpurun_out/r01_estimator_full.ncu-rep profiles/r01_estimator_full.json
"""
import csv
import json
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "smsp__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.sum",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "sm__sass_inst_executed_op_shared_ld.sum",
    "smsp__inst_executed_pipe_uniform.sum",
    # FP64 tensor pipe (DMMA) -- the counters north_star asks to be reported against peak
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "SM_C.TriageCompute.smsp__pipe_tensor_subpipe_dmma_cycles_active.avg",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.sum", "sm__inst_executed_pipe_tensor_subpipe_dmma.avg",
]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio")


def launches(src, dst):
    rows = list(csv.reader(l for l in open(src) if l.startswith('"')))
    hdr = rows[0]
    ik, iv, ig, ib = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size"), hdr.index("Block Size")
    agg = OrderedDict()
    total = 0.0
    for r in rows[1:]:
        name = r[ik].split("(")[0]
        ns = float(r[iv].replace(",", ""))
        a = agg.setdefault(name, {"n": 0, "ns": 0.0, "grid": r[ig], "block": r[ib], "max": 0.0})
        a["n"] += 1
        a["ns"] += ns
        a["max"] = max(a["max"], ns)
        total += ns
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary ({src}; --metrics gpu__time_duration.sum --clock-control none)\n\n")
        f.write(f"{len(rows) - 1} launches, {total * 1e-6:.1f} ms of GPU time (serialised, cold-cache: shares are what counts)\n\n")
        f.write("| kernel | launches | total ms | share | mean us | max us | last grid | block |\n|---|---:|---:|---:|---:|---:|---|---|\n")
        for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
            f.write(f"| {name} | {a['n']} | {a['ns'] * 1e-6:.2f} | {100 * a['ns'] / total:.1f}% | {a['ns'] / a['n'] * 1e-3:.1f} | "
                    f"{a['max'] * 1e-3:.1f} | {a['grid']} | {a['block']} |\n")
    print(open(dst).read())


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = OrderedDict()
        st = {}
        for h, u, v in zip(hdr, units, vals):
            if h == "Kernel Name":
                d["kernel"] = v.split("(")[0]
            elif h in KEYS:
                d[h] = {"value": v, "unit": u}
            else:
                m = STALL.match(h)
                if m:
                    st[m.group(1)] = float(v)
        d["stall_cycles_per_issue"] = dict(sorted(st.items(), key=lambda kv: -kv[1])[:8])
        res.append(d)
    with open(dst, "w") as f:
        json.dump({"source": src, "captures": res}, f, indent=1)
    print(json.dumps(res, indent=1)[:6000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
