#!/usr/bin/env python
"""Join an ncu launch list (gpu__time_duration.sum CSV of `bench.py --profile-step`) with the launch table
(`bench.py --profile-out`) of the same step: per plan TAG, the ncu device time (not inflated by host launch gaps).
Usage: python tools/join_ncu_tags.py launches.csv launch_table.json [tag|kernel]"""
import collections
import csv
import json
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
ncu = [(r["Kernel Name"], float(r["Metric Value"].replace(",", "")) / 1e3) for r in csv.DictReader(lines)
       if r["Metric Name"] == "gpu__time_duration.sum" and "at::" not in r["Kernel Name"]]
rows = json.load(open(sys.argv[2]))
key = sys.argv[3] if len(sys.argv) > 3 else "tag"
two = {"global_avgpool", "upsample_argmax", "confusion_matrix", "confidence"}
i = 0
agg = collections.defaultdict(lambda: [0, 0.0, 0, 0])
for r in rows:
    k = 2 if r["kernel"] in two else 1
    us = sum(t for _, t in ncu[i:i + k])
    i += k
    a = agg[r[key]]
    a[0] += 1; a[1] += us; a[2] += r["flops"]; a[3] += r["bytes"]
assert i == len(ncu), (i, len(ncu))
tot = sum(a[1] for a in agg.values())
print(f"total {tot / 1e3:.3f} ms (ncu, serialised) over {len(rows)} launches")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:34s} n={a[0]:4d} ms={a[1] / 1e3:7.3f} {100 * a[1] / tot:5.1f}%  avg_us={a[1] / a[0]:7.1f} TF/s={a[2] / a[1] / 1e6:8.1f} GB/s={a[3] / a[1] / 1e3:8.1f}")
