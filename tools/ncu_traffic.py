#!/usr/bin/env python
"""Per kernel family (the `kernel` key of bench.py's launch table): measured DRAM traffic per launch from an ncu
launch list that carries dram__bytes_read.sum / dram__bytes_write.sum / gpu__time_duration.sum for one bench step
(`bench.py --profile-step`).  Writes the JSON bench.py reads to fill `roofline.traffic`.
Usage: python tools/ncu_traffic.py launches.csv launch_table.json out.json"""
import collections
import csv
import json
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
per = collections.OrderedDict()
for r in csv.DictReader(lines):
    if "at::" in r["Kernel Name"]:
        continue
    d = per.setdefault(r["ID"], {"name": r["Kernel Name"]})
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
    d[r["Metric Name"]] = v * scale
ncu = list(per.values())
rows = json.load(open(sys.argv[2]))
two = {"global_avgpool", "upsample_argmax", "confusion_matrix", "confidence"}
agg = collections.defaultdict(lambda: dict(launches=0, dram_bytes=0.0, us=0.0, algorithmic_bytes=0, flops=0))
inst = collections.defaultdict(lambda: dict(launches=0, dram_bytes=0.0, us=0.0, algorithmic_bytes=0, flops=0))
i = 0
for r in rows:
    k = 2 if r["kernel"] in two else 1
    # kernel family and kernel INSTANCE (same tag + same algorithmic work = same shapes; bench.py's roofline key)
    for a in (agg[r["kernel"]], inst[f'{r["tag"]}|{r["flops"]}|{r["bytes"]}']):
        a["launches"] += 1
        for d in ncu[i:i + k]:
            a["dram_bytes"] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
            a["us"] += d.get("gpu__time_duration.sum", 0.0)
        a["algorithmic_bytes"] += r["bytes"]; a["flops"] += r["flops"]
    i += k
assert i == len(ncu), (i, len(ncu))
out = {"source": sys.argv[1], "note": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum over one "
       "bench step (serialised, cold caches); traffic_per_launch = (read + write) / launches", "kernels": {}}
for k, a in agg.items():
    out["kernels"][k] = dict(launches=a["launches"], traffic_per_launch=a["dram_bytes"] / a["launches"],
                             algorithmic_bytes_per_launch=a["algorithmic_bytes"] / a["launches"],
                             flops_per_launch=a["flops"] / a["launches"], ncu_us_per_launch=a["us"] / a["launches"])
out["instances"] = {k: dict(launches=a["launches"], traffic_per_launch=a["dram_bytes"] / a["launches"],
                            algorithmic_bytes_per_launch=a["algorithmic_bytes"] / a["launches"],
                            flops_per_launch=a["flops"] / a["launches"], ncu_us_per_launch=a["us"] / a["launches"])
                    for k, a in inst.items()}
json.dump(out, open(sys.argv[3], "w"), indent=1)
for k, v in sorted(out["kernels"].items(), key=lambda kv: -kv[1]["ncu_us_per_launch"] * kv[1]["launches"]):
    print(f"{k:22s} n={v['launches']:4d} dram/launch={v['traffic_per_launch'] / 1e6:9.2f} MB  algorithmic={v['algorithmic_bytes_per_launch'] / 1e6:9.2f} MB  ratio={v['traffic_per_launch'] / max(v['algorithmic_bytes_per_launch'], 1):5.2f}")
