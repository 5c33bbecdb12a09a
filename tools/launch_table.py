#!/usr/bin/env python
"""Aggregate a bench.py --profile-out launch table by tag (or kernel): count, ms, share, TFLOP/s, GB/s."""
import collections
import json
import sys

rows = json.load(open(sys.argv[1]))
key = sys.argv[2] if len(sys.argv) > 2 else "tag"
agg = collections.defaultdict(lambda: [0, 0.0, 0, 0])
for r in rows:
    a = agg[r[key]]
    a[0] += 1; a[1] += r["ms"]; a[2] += r["flops"]; a[3] += r["bytes"]
tot = sum(a[1] for a in agg.values())
print(f"total {tot:.3f} ms over {len(rows)} launches")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:36s} n={a[0]:4d} ms={a[1]:7.3f} {100 * a[1] / tot:5.1f}%  avg_us={1e3 * a[1] / a[0]:7.1f} TF/s={a[2] / a[1] / 1e9:8.1f} GB/s={a[3] / a[1] / 1e6:8.1f}")
