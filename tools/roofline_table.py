#!/usr/bin/env python
"""Per-tag roofline table of one bench step from the per-launch CUDA-event table bench.py writes (--profile-out):
for every launch the roofline time is max(algorithmic FLOPs / bf16 peak, algorithmic bytes / HBM peak) (SURVEY 8d work
model, MEASURED_PEAKS.json peaks: burst bf16 for a launch timed alone); per tag: launches, measured time, roofline time, the
bound that dominates it and the fraction of the roofline achieved.  No GPU needed.
Usage: python tools/roofline_table.py profiles/r4e_launch_table.json > profiles/r4e_roofline_by_tag.md"""
import collections
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
rows = json.loads(Path(sys.argv[1]).read_text())
pk = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
hbm, tc = pk.get("hbm_gbs", 6545.3) * 1e9, pk.get("bf16_tflops", 1605.1) * 1e12
agg = collections.OrderedDict()
for r in rows:
    t_f, t_b = r["flops"] / tc, r["bytes"] / hbm
    d = agg.setdefault(r["tag"], dict(n=0, ms=0.0, roof=0.0, tf=0.0, tb=0.0, flops=0, bytes=0, kernel=r["kernel"]))
    d["n"] += 1; d["ms"] += r["ms"]; d["roof"] += max(t_f, t_b) * 1e3; d["tf"] += t_f * 1e3; d["tb"] += t_b * 1e3
    d["flops"] += r["flops"]; d["bytes"] += r["bytes"]
tot_ms = sum(d["ms"] for d in agg.values()); tot_roof = sum(d["roof"] for d in agg.values())
print(f"# Roofline by tag: {sys.argv[1]}\n")
print(f"{len(rows)} launches, {tot_ms:.3f} ms of per-launch CUDA-event time (each launch timed alone; the replayed graph overlaps them: "
      f"see ms_per_step of the bench line); sum of per-launch roofline times {tot_roof:.3f} ms = {tot_roof / tot_ms:.2f} of it.  "
      f"Peaks: HBM {hbm / 1e9:.0f} GB/s, bf16 {tc / 1e12:.0f} TFLOP/s (MEASURED_PEAKS.json, burst).\n")
print("| tag | kernel | launches | measured ms | share | roofline ms | bound | fraction of roofline | TFLOP/s | GB/s |")
print("|---|---|---:|---:|---:|---:|---|---:|---:|---:|")
for tag, d in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
    bound = "tensor" if d["tf"] > d["tb"] else "hbm"
    print(f"| `{tag}` | `{d['kernel']}` | {d['n']} | {d['ms']:.3f} | {100 * d['ms'] / tot_ms:.1f}% | {d['roof']:.3f} | {bound} | "
          f"{d['roof'] / max(d['ms'], 1e-12):.2f} | {d['flops'] / max(d['ms'], 1e-12) / 1e9:.0f} | {d['bytes'] / max(d['ms'], 1e-12) / 1e6:.0f} |")
print("\nNotes: the byte model of a bilinear DOWN-scale counts the whole source map (SURVEY 8d: (P_in + P_out) C e) although a 4x "
      "down-scale only touches the sampled source pixels, hence the fraction above 1 on `Cell.resize_pp`; `ASPP.pool_bias`, `EDM.mlp` "
      "and the gathers are latency-sized launches (a few KB), their roofline time is ~0.")
