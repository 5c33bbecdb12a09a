#!/usr/bin/env python
"""Which rounding point of the bf16 path costs how much (no GPU needed): the launch-layer stand-in of tests/sim_backend.py
run in bf16 with individual rounding points switched off — bf16 weights, the bf16 SepConv depthwise tile — next to the full
storage model, against the fp32 oracle, on the input and weights of tools/bf16_parity_probe.py (257 x 513, seed 4321).
Activation storage is always bf16 (that is what a bf16 path is).  Usage: python tools/bf16_storage_bisect.py"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / 'tests')); sys.path.insert(0, str(ROOT))
import torch, add_b200, sim_backend
import torch.nn.functional as F
from oracle import add_oracle as orc
class MP:
    def setattr(self,o,n,v): setattr(o,n,v)
sim_backend.install(MP())
MODE = {"w": True, "dw": True, "act": True}
def run(mode):
    MODE.update(mode)
    # weights / dw rounding toggles via monkeypatching torch.bfloat16 casts inside sim: simplest = wrap conv/sepconv
    def conv(self, x, y, cw, stride=1, pad=0, dil=1, flags=0, tag="conv", image_bias=None):
        if x.relud: flags &= ~sim_backend.RELU_IN
        def r():
            xin = sim_backend._window(sim_backend._x(x, flags & 1), y.h, y.w, cw.kh, cw.kw, stride, pad, dil)
            w = cw.w_h.permute(3,2,0,1).contiguous()
            if x.dtype == torch.bfloat16 and MODE["w"]: w = w.to(torch.bfloat16).float()
            out = F.conv2d(xin, w, None, stride, 0, dil)
            if image_bias is not None: out = out + image_bias.float().view(x.n, cw.cout,1,1)
            elif cw.bias is not None: out = out + cw.bias.float().view(1,-1,1,1)
            sim_backend._store(y, out, flags)
        sim_backend._do(self, r, tag, "conv2d")
    def sep(self, x, y, w_dw, pw, k, flags, tag="sephalf"):
        if x.relud and (flags & 1): flags &= ~1
        def r():
            xin = sim_backend._x(x, flags & 1)
            d = F.conv2d(xin, w_dw.float().permute(2,0,1).unsqueeze(1).contiguous(), None, 1, k//2, 1, x.c)
            w = pw.w_h.permute(3,2,0,1).contiguous()
            if x.dtype == torch.bfloat16:
                if MODE["dw"]: d = d.to(torch.bfloat16).float()
                if MODE["w"]: w = w.to(torch.bfloat16).float()
            out = F.conv2d(d, w)
            if pw.bias is not None: out = out + pw.bias.float().view(1,-1,1,1)
            sim_backend._store(y, out, flags)
        sim_backend._do(self, r, tag, "sepconv_half")
    from add_b200.runtime import Builder
    Builder.conv = conv; Builder.sepconv_half = sep
na, ci, low = add_b200.NETWORKS["searched-dense"][2]
arch = orc.Arch(na, ci, low_level_layer=low)
x, _ = orc.synthetic_batch(1, 257, 513, seed=4321)
for wname in ("randomized", "calibrated"):
    base = add_b200.build_add("searched-dense", 2, 20, seed=1)
    sd = {k: v.clone() for k, v in base.state_dict().items()}
    sd = orc.randomize_bn_(sd, 21) if wname == "randomized" else orc.calibrate_bn_(sd, arch)
    with torch.no_grad(): ref = orc.add_forward(sd, arch, x)
    for label, mode in (("all roundings", dict(w=True,dw=True)), ("activations only (fp32 weights, fp32 dw tile)", dict(w=False,dw=False)),
                        ("activations + weights", dict(w=True,dw=False)), ("activations + dw tile", dict(w=False,dw=True))):
        run(mode)
        net = add_b200.build_add("searched-dense", 2, 20, seed=1); net.load_state_dict(sd); net.eval(); net.set_precision("bf16")
        outs = net(x)
        res = []
        for o, r in zip(outs, ref):
            o, r = o.double(), r.double(); sc = r.abs().max()
            res.append('rms %.4f agree %.4f' % (float((o-r).pow(2).mean().sqrt()/sc), float((o.argmax(1)==r.argmax(1)).float().mean())))
        print(wname, '|', label, '|', ' ; '.join(res))
