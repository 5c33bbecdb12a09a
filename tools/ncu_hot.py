#!/usr/bin/env python
"""Hot SASS instructions of one kernel out of `ncu -i rep --page source --csv` (first section of the file).
Usage: python tools/ncu_hot.py src.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
hdr = rows[1]
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) == len(hdr):
        data.append(r)
ia, isamp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
tot, tots = sum(int(r[ia]) for r in data), sum(int(r[isamp]) for r in data)
print(rows[0][1][:100])
print("warp instructions", tot, "samples", tots, "sass lines", len(data))
print("--- by executed count")
for r in sorted(data, key=lambda r: -int(r[ia]))[:n]:
    print(f"{int(r[ia]):9d} {100 * int(r[ia]) / tot:5.1f}%  samp {int(r[isamp]):6d}  {r[isrc].strip()[:80]}")
print("--- by stall samples")
for r in sorted(data, key=lambda r: -int(r[isamp]))[:n]:
    print(f"{int(r[ia]):9d}  samp {int(r[isamp]):6d} {100 * int(r[isamp]) / max(tots, 1):5.1f}%  {r[isrc].strip()[:80]}")
