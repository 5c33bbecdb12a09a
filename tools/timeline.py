#!/usr/bin/env python
"""Kernel timeline of one bench step (BASELINE config 2) from CUPTI via torch.profiler: per kernel start / duration /
stream of the CUDA-graph replays, then the union busy time, idle gaps and the concurrency profile.  Diagnostic only
(the numbers are taken under a profiler — never a bench value).  Usage: python tools/timeline.py [out.json]"""
import json
import sys
from pathlib import Path

import torch
from torch.profiler import profile, ProfilerActivity

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import add_b200  # noqa: E402

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/timeline.json"
dev = torch.device("cuda:0")
net = add_b200.build_add("searched-dense", 2, 20, seed=1).to(dev)
net.set_precision("bf16")
net.use_cuda_graph = True
torch.manual_seed(203)
edm = add_b200.EDM().eval().to(dev)
x, gt = add_b200.synthetic_batch(8, 1024, 2048, seed=1234)
x, gt = x.to(dev), gt.to(dev)
_, _, confs = net.dynamic_evaluate(x, gt, -1e30, edm)
vals = sorted(float(c) for c in confs)
thr = 0.5 * (vals[3] + vals[4])
for _ in range(5):
    net.dynamic_evaluate(x, gt, thr, edm)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        net.dynamic_evaluate(x, gt, thr, edm)
    torch.cuda.synchronize()
evs = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        evs.append(dict(name=e.name[:90], start=e.time_range.start, dur=e.time_range.end - e.time_range.start))
evs.sort(key=lambda d: d["start"])
if not evs:
    print("no CUDA events recorded")
    sys.exit(1)
t0 = evs[0]["start"]
for e in evs:
    e["start"] -= t0
# the last step only: events after the 2/3 mark gap — find step boundaries as the largest idle gaps
Path(out).write_text(json.dumps(evs))
# union busy time and concurrency
pts = []
for e in evs:
    pts.append((e["start"], 1)); pts.append((e["start"] + e["dur"], -1))
pts.sort()
busy = 0.0; conc_time = {}; cur = 0; last = pts[0][0]
for t, d in pts:
    if cur > 0:
        busy += t - last
    conc_time[cur] = conc_time.get(cur, 0.0) + (t - last)
    cur += d; last = t
span = evs[-1]["start"] + evs[-1]["dur"]
print(f"kernels {len(evs)}  span {span / 1e3:.3f} ms  busy(union) {busy / 1e3:.3f} ms  idle {100 * (1 - busy / span):.1f} %")
print("time at concurrency level (ms):", {k: round(v / 1e3, 3) for k, v in sorted(conc_time.items())})
agg = {}
for e in evs:
    k = e["name"].split("(")[0][:60]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += e["dur"]
for k, (n, d) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
    print(f"{d / 1e3:9.3f} ms  n={n:5d}  avg={d / n:8.1f} us  {k}")
