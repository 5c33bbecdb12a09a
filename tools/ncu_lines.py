#!/usr/bin/env python
"""Per CUDA source line totals (instructions executed, stall samples) from
`ncu -i rep --page source --csv --print-source cuda,sass` output.  Usage: python tools/ncu_lines.py file.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cur_file = ""
agg = {}
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    if r[0] != "":      # a CUDA source line row: totals for the line
        try:
            key = (cur_file, int(r[0]), r[1].strip()[:70])
            a = agg.setdefault(key, [0, 0])
            a[0] += int(r[ia]); a[1] += int(r[isamp])
        except ValueError:
            pass
tot = sum(a[0] for a in agg.values()); tots = sum(a[1] for a in agg.values())
print("total warp instructions", tot, "samples", tots)
for (f, ln, src), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    print(f"{a[0]:9d} {100 * a[0] / max(tot, 1):5.1f}%  samp {a[1]:5d} {100 * a[1] / max(tots, 1):5.1f}%  {f}:{ln}  {src}")
