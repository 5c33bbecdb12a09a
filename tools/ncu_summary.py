#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name, launches,
total device time and share.  Usage: python tools/ncu_summary.py launches.csv > profiles/xxx.md"""
import csv
import re
import sys
from collections import defaultdict


def short(name: str) -> str:
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    m = re.match(r"(?:void\s+)?([A-Za-z_0-9:]+)", name)
    base = m.group(1) if m else name
    targs = re.search(r"<(.*)>", name)
    return base + ("<" + targs.group(1)[:60] + ">" if targs else "")


def main(path):
    rows = []
    with open(path) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            unit = r.get("Metric Unit", "ns")
            scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
            rows.append((short(r["Kernel Name"]), v * scale, r["Grid Size"], r["Block Size"]))
    agg = defaultdict(lambda: [0, 0.0])
    for k, us, *_ in rows:
        agg[k][0] += 1
        agg[k][1] += us
    total = sum(v[1] for v in agg.values())
    print(f"# ncu launch list summary: {path}\n")
    print(f"{len(rows)} launches, {total / 1e3:.3f} ms total device time (cold-cache, serialised: compare SHARES)\n")
    print("| kernel | launches | total ms | share | avg us |")
    print("|---|---:|---:|---:|---:|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {us / 1e3:.3f} | {100 * us / total:.1f}% | {us / n:.1f} |")


if __name__ == "__main__":
    main(sys.argv[1])
