#!/usr/bin/env python
"""Summarise an `ncu --metrics ... --csv` launch log per kernel: launches, mean duration, DRAM bytes, warp instructions,
issue utilisation, active lanes, registers, FP32 thread-ops.   usage: ncu_kernel_table.py log.csv [out.csv]"""
import csv, sys, re
from collections import defaultdict
rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[hi]
data = defaultdict(lambda: defaultdict(list))
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    d = dict(zip(hdr, r))
    name = re.sub(r"\(.*", "", d["Kernel Name"]).replace("void ", "").replace("jr::", "")
    key = (name, d.get("Grid Size", ""), d.get("Block Size", ""))
    try:
        data[key][d["Metric Name"]].append((float(d["Metric Value"].replace(",", "")), d["Metric Unit"]))
    except ValueError:
        pass
def mean(key, m, scale=None):
    v = data[key].get(m)
    if not v: return float("nan")
    x = sum(a for a, _ in v) / len(v)
    u = v[0][1]
    if scale == "us": x *= {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}.get(u, 1)
    if scale == "MB": x *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1, "Gbyte": 1e3}.get(u, 1)
    return x
out = [("kernel", "grid", "block", "launches", "us", "dram_MB", "GB/s", "frac_of_6454", "warp_inst_M", "issue_%", "warps_active_%", "lanes", "regs", "fp32_Gop")]
tot = 0
for key in data:
    n = len(data[key]["gpu__time_duration.sum"])
    us = mean(key, "gpu__time_duration.sum", "us")
    mb = mean(key, "dram__bytes_read.sum", "MB") + mean(key, "dram__bytes_write.sum", "MB")
    fl = (mean(key, "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum") + mean(key, "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum")
          + 2 * mean(key, "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum")) / 1e9
    gbs = mb / us * 1e3 if us == us and us > 0 else float("nan")
    out.append((key[0], key[1], key[2], n, round(us, 1), round(mb, 1), round(gbs, 0), round(gbs / 6453.7, 3),
                round(mean(key, "smsp__inst_executed.sum") / 1e6, 2), round(mean(key, "smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
                round(mean(key, "sm__warps_active.avg.pct_of_peak_sustained_active"), 1),
                round(mean(key, "smsp__thread_inst_executed_per_inst_executed.ratio"), 1), int(mean(key, "launch__registers_per_thread") or 0)
                if mean(key, "launch__registers_per_thread") == mean(key, "launch__registers_per_thread") else "", round(fl, 2)))
    tot += us * n
w = csv.writer(open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout)
for r in out: w.writerow(r)
print(f"# total kernel time of the captured launches: {tot:.0f} us", file=sys.stderr)
