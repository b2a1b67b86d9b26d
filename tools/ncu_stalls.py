#!/usr/bin/env python
"""Stall-reason samples of an ncu source-page export, total and per line range.
usage: ncu_stalls.py src.csv  name:file:lo-hi ..."""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
fname = ''; hdr = None
per = defaultdict(lambda: defaultdict(int))
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit(): continue
    for name, v in zip(hdr, r):
        if name.startswith('stall_') and '(Not Issued)' not in name and v.strip().isdigit():
            per[(fname, int(r[0]))][name] += int(v)
def show(title, keys):
    tot = defaultdict(int)
    for k in keys:
        for n, v in per[k].items(): tot[n] += v
    s = sum(tot.values()) or 1
    top = sorted(tot.items(), key=lambda kv: -kv[1])[:7]
    print(f"{title:18s} {s:7d}  " + "  ".join(f"{n[6:]}:{100*v/s:.0f}%" for n, v in top))
show('ALL', list(per.keys()))
for spec in sys.argv[2:]:
    name, f, rng = spec.split(':'); lo, hi = map(int, rng.split('-'))
    show(name, [k for k in per if k[0] == f and lo <= k[1] <= hi])
