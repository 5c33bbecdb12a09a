import sys, torch, time
sys.path.insert(0, '/root/repo')
import __graft_entry__ as g; g.build()
import add_b200
dev = torch.device('cuda:0')
net = add_b200.build_add("searched-dense", 2, 20, seed=1).to(dev); net.set_precision("bf16"); net.use_cuda_graph = True
torch.manual_seed(203); edm = add_b200.EDM().eval().to(dev)
x, gt = add_b200.synthetic_batch(8, 1024, 2048, seed=1234); xd, gd = x.to(dev), gt.to(dev)
_, _, confs = net.dynamic_evaluate(xd, gd, -1e30, edm)
vals = sorted(float(c) for c in confs); thr = 0.5 * (vals[3] + vals[4])
for _ in range(3): net.dynamic_evaluate(xd, gd, thr, edm)
def timed(fn, n=20):
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n
t_full = timed(lambda: net.dynamic_evaluate(xd, gd, thr, edm))
runner = next(v for k, v in net._plans.items() if k[0] == "edm" and k[5] == "evaluate")
plans = list(runner.last_plans)
def replay():
    for p in plans: p.run()
t_replay = timed(replay)
print(f"full step {t_full:.3f} ms   graphs replayed back-to-back (no host gate) {t_replay:.3f} ms   host/gate overhead {t_full - t_replay:.3f} ms")
