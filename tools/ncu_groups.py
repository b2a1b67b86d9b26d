#!/usr/bin/env python
"""Group an ncu source-page export (see ncu_source_lines.py) by line ranges given in a spec file / inline.
usage: ncu_groups.py src.csv images  name:file:lo-hi ..."""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
images = float(sys.argv[2])
agg = defaultdict(lambda: [0, 0, 0]); fname = ''; hdr = None
for r in rows:
    if not r: continue
    if r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name': continue
    if r[0] == 'Line No': hdr = r; continue
    if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit(): continue
    d = dict(zip(hdr, r))
    num = lambda v: int(v) if v.strip().lstrip('-').isdigit() else 0
    a = agg[(fname, int(r[0]))]
    a[0] += num(d['Instructions Executed']); a[1] += num(d['Thread Instructions Executed']); a[2] += num(d['# Samples'])
tot = sum(a[0] for a in agg.values()); tots = sum(a[2] for a in agg.values()); used = set()
print(f"total {tot}  per image {tot/images:.0f}  samples {tots}")
for spec in sys.argv[3:]:
    name, f, rng = spec.split(':'); lo, hi = map(int, rng.split('-'))
    s = [0, 0, 0]
    for (ff, l), a in agg.items():
        if ff == f and lo <= l <= hi and (ff, l) not in used:
            s[0] += a[0]; s[1] += a[1]; s[2] += a[2]; used.add((ff, l))
    print(f"{name:22s} {100*s[0]/tot:6.2f}%  per-image {s[0]/images:8.0f}  lanes {s[1]/max(s[0],1):5.1f}  samples {100*s[2]/tots:5.1f}%")
rest = defaultdict(lambda: [0, 0, 0])
for k, a in agg.items():
    if k in used: continue
    b = rest[k[0]]; b[0] += a[0]; b[1] += a[1]; b[2] += a[2]
for f, s in rest.items():
    print(f"(rest) {f:22s} {100*s[0]/tot:6.2f}%  per-image {s[0]/images:8.0f}  lanes {s[1]/max(s[0],1):5.1f}  samples {100*s[2]/tots:5.1f}%")
