#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per source line.

  ncu -i prof.ncu-rep --page source --csv --print-source cuda,sass > src.csv
  python tools/ncu_source_lines.py src.csv [top_n]

Prints warp instructions, share, stall samples and active lanes per (file, line)."""
import csv
import sys
from collections import defaultdict


def main() -> None:
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    rows = list(csv.reader(open(path)))
    agg = defaultdict(lambda: [0, 0, 0, ""])  # inst, thread inst, samples, source
    fname = ""
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or not r[0].strip().isdigit():
            continue   # SASS rows (empty line number) are already summed into their source-line row
        d = dict(zip(hdr, r))
        # two "Source" columns: the dict keeps the second (SASS); the first is r[1]
        try:
            inst = int(d["Instructions Executed"]); tinst = int(d["Thread Instructions Executed"])
            samp = int(d["# Samples"])
        except (ValueError, KeyError):
            continue
        k = (fname, int(r[0]))
        a = agg[k]
        a[0] += inst; a[1] += tinst; a[2] += samp
        if not a[3]:
            a[3] = r[1].strip()
    tot = sum(a[0] for a in agg.values()) or 1
    tots = sum(a[2] for a in agg.values()) or 1
    print(f"total warp instructions {tot}, samples {tots}")
    print("file:line                 inst      %inst  %samp  lanes  source")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        lanes = a[1] / a[0] if a[0] else 0
        print(f"{k[0][:18]:18s}:{k[1]:<5d} {a[0]:10d} {100*a[0]/tot:6.2f} {100*a[2]/tots:6.2f} {lanes:6.1f}  {a[3][:90]}")


if __name__ == "__main__":
    main()
