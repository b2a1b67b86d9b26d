#!/usr/bin/env python
"""profiles/roofline_traffic.json from an `ncu --set full` capture of the bench's dominant kernel.

  python tools/update_roofline_traffic.py gpurun_out/prof.ncu-rep

Records DRAM bytes read + written and warp instructions per launch, and the hash of the kernel sources the capture
was taken on (bench.py reports `roofline.traffic` only while that hash matches: a stale capture is not a number)."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main() -> None:
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, vals = rows[0], rows[2]
    d = dict(zip(hdr, vals))
    rd, wr = float(d["dram__bytes_read.sum"]), float(d["dram__bytes_write.sum"])
    unit = dict(zip(hdr, rows[1]))
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd *= scale.get(unit["dram__bytes_read.sum"], 1)
    wr *= scale.get(unit["dram__bytes_write.sum"], 1)
    inst = float(d["smsp__inst_executed.sum"])
    from bench import _csrc_sha16

    out = {
        "kernel": d.get("Kernel Name", ""),
        "dram_bytes_per_launch": rd + wr,
        "dram_bytes_read": rd, "dram_bytes_write": wr,
        "warp_instructions_per_launch": inst,
        "duration_us_under_ncu": float(d.get("gpu__time_duration.sum", 0)) / (1e3 if unit.get("gpu__time_duration.sum") == "ns" else 1),
        "csrc_sha16": _csrc_sha16(),
        "source": f"ncu --set full --clock-control none of {d.get('Kernel Name', '')} on the bench workload "
                  f"({os.path.basename(rep)}): dram__bytes_read.sum + dram__bytes_write.sum, smsp__inst_executed.sum",
    }
    json.dump(out, open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
