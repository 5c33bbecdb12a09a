#!/usr/bin/env python
"""bf16 (tensor-core path) vs the fp32 CPU oracle, whole network, per exit: max-norm relative error, argmax agreement
overall and on the pixels whose fp32 top-2 logit gap exceeds a multiple of the measured error, on BN-randomised and on
BN-calibrated weights, at several sizes (last = BASELINE config 2, 1024x2048).  Also the EDM gate value in bf16 / fp32.
Usage: python tools/bf16_parity_probe.py [--sizes 257x513,1024x2048] [--json out.json]"""
import argparse
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import add_b200  # noqa: E402
from oracle import add_oracle as orc  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="129x257,257x513,1024x2048")
ap.add_argument("--json", default=None)
a = ap.parse_args()
sizes = [tuple(int(v) for v in s.split("x")) for s in a.sizes.split(",")]
dev = torch.device("cuda:0")
torch.set_num_threads(max(1, torch.get_num_threads()))
na, ci, low = add_b200.NETWORKS["searched-dense"][2]
arch = orc.Arch(na, ci, low_level_layer=low)
rows = []


def metrics(o, r):
    o, r = o.double().cpu(), r.double()
    err = float((o - r).abs().max() / r.abs().max())
    rms = float((o - r).pow(2).mean().sqrt() / r.abs().max())
    agree = (o.argmax(1) == r.argmax(1))
    top2 = r.topk(2, 1).values
    gap = (top2[:, 0] - top2[:, 1]) / r.abs().max()
    out = dict(rel_err=err, rms_rel=rms, argmax_agree=float(agree.float().mean()))
    for t in (1e-3, 1e-2, 3e-2, 5e-2):
        m = gap > t
        out[f"agree_gap>{t:g}"] = float(agree[m].float().mean()) if m.any() else None
        out[f"frac_gap>{t:g}"] = float(m.float().mean())
    return out


for wname in ("randomized_bn", "calibrated_bn"):
    net = add_b200.build_add("searched-dense", 2, 20, seed=1)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    sd = orc.randomize_bn_(sd, 21) if wname == "randomized_bn" else orc.calibrate_bn_(sd, arch)
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    torch.manual_seed(203)
    edm = add_b200.EDM().eval()
    edm_sd = {k: v.detach().clone() for k, v in edm.state_dict().items()}
    edm = edm.to(dev)
    for (h, w) in sizes:
        x, gt = orc.synthetic_batch(1, h, w, seed=4321)
        t0 = time.time()
        with torch.no_grad():
            ref = orc.add_forward(sd, arch, x)
            _, feat = orc.add_get_feature(sd, arch, x)
            conf_ref = float(orc.edm_forward(edm_sd, feat.clone()))
            ref_dyn, _, _ = orc.add_dynamic_inference(sd, arch, x, 1e30, 'edm', edm_sd)
        t_cpu = time.time() - t0
        # stock PyTorch bf16 (cuDNN / ATen kernels, bf16 weights + activations, channels_last) on the same weights: the
        # precision a user of the reference gets from `model.to(torch.bfloat16)`; more roundings than our fused path
        def cast(v):
            v = v.to(dev)
            if v.is_floating_point():
                v = v.to(torch.bfloat16)
                if v.dim() == 4:
                    v = v.contiguous(memory_format=torch.channels_last)
            return v
        with torch.no_grad():
            sdd = {k: cast(v) for k, v in sd.items()}
            t_outs = orc.add_forward(sdd, arch, cast(x))
            _, t_feat = orc.add_get_feature(sdd, arch, cast(x))
        for e, (o, r) in enumerate(zip(t_outs, ref)):
            m = metrics(o.float(), r)
            m.update(weights=wname, size=f"{h}x{w}", precision="torch_bf16", exit=e, edm_value=0.0, edm_ref=conf_ref,
                     max_abs_ref=float(r.abs().max()), cpu_s=t_cpu,
                     feat_rms_rel=float(((t_feat.float().cpu().double() - feat.double()).pow(2).mean().sqrt() / feat.double().abs().max())))
            rows.append(m)
            print(json.dumps(m), flush=True)
        for prec in ("fp32", "bf16"):
            net.set_precision(prec)
            _, f_ours = net.get_feature(x.to(dev))
            feat_err = float(((f_ours.cpu().double() - feat.double()).pow(2).mean().sqrt() / feat.double().abs().max()))
            outs = net(x.to(dev))
            _, _, _, cv = net.dynamic_inference(x.to(dev), threshold=-1e30, confidence='edm', edm=edm)
            y_dyn, _, _, _ = net.dynamic_inference(x.to(dev), threshold=1e30, confidence='edm', edm=edm)   # early exit taken
            outs = list(outs) + [y_dyn]
            for e, (o, r) in enumerate(zip(outs, ref + [ref_dyn])):
                m = metrics(o, r)
                m.update(weights=wname, size=f"{h}x{w}", precision=prec, exit=e, edm_value=float(cv), edm_ref=conf_ref,
                         max_abs_ref=float(r.abs().max()), cpu_s=t_cpu, feat_rms_rel=feat_err)
                rows.append(m)
                print(json.dumps(m), flush=True)
if a.json:
    Path(a.json).write_text(json.dumps(rows, indent=1))
