#!/usr/bin/env python
"""How much of a bench step is the gate's host round trip?  Times (a) the normal step (`dynamic_evaluate`: trunk graph
-> D2H of N gate values -> host decision -> head / remaining-trunk graphs) and (b) the SAME graphs replayed back to
back with the indices left as the last step wrote them (no host decision in between).  (a) - (b) = GPU idle time per
step caused by the gate.  Usage: python tools/gate_bubble.py [steps]"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import add_b200  # noqa: E402
from add_b200 import dynamic  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
net = add_b200.build_add("searched-dense", 2, 20, seed=1).to(dev)
net.set_precision("bf16")
net.use_cuda_graph = True
torch.manual_seed(203)
edm = add_b200.EDM().eval().to(dev)
x_host, gt_host = add_b200.synthetic_batch(8, 1024, 2048, seed=1234, pin=True)
x, gt = x_host.to(dev), gt_host.to(dev)
_, _, confs = net.dynamic_evaluate(x, gt, -1e30, edm, "reference")
vals = sorted(float(c) for c in confs)
thr = 0.5 * (vals[3] + vals[4])


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, 1e3 * (time.perf_counter() - t0) / steps


a_gpu, a_wall = timed(lambda: net.dynamic_evaluate(x, gt, thr, edm, "reference"))
r = dynamic._get_runner(net, x, edm, "evaluate", "reference", gt, False)
plans = list(r.last_plans)


def replay():
    for p in plans:
        p.run()


b_gpu, b_wall = timed(replay)
print(f"normal step       : {a_gpu:.3f} ms (device events), {a_wall:.3f} ms wall")
print(f"graphs back to back: {b_gpu:.3f} ms (device events), {b_wall:.3f} ms wall   [{len(plans)} graphs, one stream]")
print(f"gate round trip    : {a_gpu - b_gpu:+.3f} ms per step")
