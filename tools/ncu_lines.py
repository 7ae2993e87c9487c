#!/usr/bin/env python
"""Per-source-line executed warp-instructions (per 32-pixel group) of an .ncu-rep captured with
--import-source on.  usage: tools/ncu_lines.py rep [npix] [min_per_group]"""
import csv, subprocess, sys
rep = sys.argv[1]
npix = int(sys.argv[2]) if len(sys.argv) > 2 else 4096 * 4096
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
cur_file = ""
tot = 0.0
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    if r and r[0] == "Line No":
        hdr = r
        ie = hdr.index("Instructions Executed")
        ss = hdr.index("# Samples")
        continue
    if hdr and len(r) > ie and r[0].isdigit() and r[ie].isdigit():
        n = int(r[ie]) / (npix / 32)
        tot += n
        if n >= thr:
            print(f"{cur_file}:{r[0]:>5s} {n:8.1f} inst/group  samples {r[ss]:>6s}  {r[1].strip()[:90]}")
print(f"total {tot:.1f}")
