"""Early-exit inference (reference ADD.py:379-488) as segmented launch plans.

The reference decides per image on the host (`if confidence_value > threshold`, a device sync).
Here the network is recorded once per input shape as segments — trunk up to each exit (+ the EDM
gate), one early-exit head per exit, the final head — and the host replays only the segments the
gate selects.  Quirks reproduced (SURVEY Q3–Q5): the early exit uses the 2^-L `aspp_size` (the
feature is bilinearly up-sampled ×4 before ASPP), EDM's in-place ReLU is visible to the exit's
resize and to later cells, the last exit never resizes, EDM exits when value <= threshold.
"""
from __future__ import annotations

from typing import Dict, List

import torch

from . import runtime as rt
from .runtime import Builder, Plan, View, RELU_IN
from .operations import _confidence


class _DynPlan:
    def __init__(self, net, shape, device, precision: str, confidence: str, edm):
        self.generation = rt.generation()
        n, _, H, W = shape
        assert n == 1
        nc = net._num_classes
        b = Builder(device, rt.act_dtype(precision), record=True)
        self.builder = b
        self.x_static = b.raw(shape, torch.float32)
        aspp_size = net._aspp_size((H, W), net.network_arch[-1])          # ADD.py:383-384
        st: dict = {}
        self.trunks: List[Plan] = []
        self.heads: List[Plan] = []
        self.conf: List[torch.Tensor] = []
        self.outs: List[torch.Tensor] = []
        exits = [i for i in range(net.num_net) if i in net.C_index and i != net.num_net - 1]
        self.exits = exits
        done = -1
        for k, i in enumerate(exits):
            start = len(b.launches)
            net._emit_trunk(b, self.x_static, done + 1, i, st)
            done = i
            y = net._feature(st, i)
            relu_feature = False
            if confidence == 'edm':
                self.conf.append(edm.emit_edm(b, y))
                # EDM.forward's in-place ReLU (ADD.py:516-519) mutates the feature every later reader sees
                yr = b.alloc(y.n, y.h, y.w, y.c)
                b.bilinear(y, yr, RELU_IN, "EDM.inplace_relu")
                if i > 2:
                    st["cur"] = yr
                else:
                    st["two"][1] = yr
                y = yr
                relu_feature = True
            self.trunks.append(Plan(b, start, len(b.launches)))
            # early-exit head k (conv_aspp_iter == k: every earlier exit was skipped, ADD.py:422)
            start = len(b.launches)
            if confidence != 'edm' and not (y.h < aspp_size[0] or y.w < aspp_size[1]):
                raise NotImplementedError("non-EDM gate with a feature not smaller than aspp_size: the reference "
                                          "skips the head and scores the raw feature map (ADD.py:465-476)")
            logits = net._emit_exit_lowres(b, y, st, i, aspp_size, k, True, relu_feature)
            out = b.raw((1, nc, H, W), torch.float32)
            b.upsample_logits(logits, out, H, W, "ADD.upsample_logits")
            self.outs.append(out)
            self.heads.append(Plan(b, start, len(b.launches)))
        # remaining cells + last exit (never resized in the EDM path: ADD.py:433-435)
        start = len(b.launches)
        last = net.num_net - 1
        net._emit_trunk(b, self.x_static, done + 1, last, st)
        y = net._feature(st, last)
        logits = net._emit_exit_lowres(b, y, st, last, aspp_size, 0, resize=(confidence != 'edm'))
        out = b.raw((1, nc, H, W), torch.float32)
        b.upsample_logits(logits, out, H, W, "ADD.upsample_logits")
        self.outs.append(out)
        self.tail = Plan(b, start, len(b.launches))
        if net.use_cuda_graph:
            for p in self.trunks + self.heads + [self.tail]:
                p.capture()

    @property
    def n_launches(self):
        return len(self.builder.launches)


def _get_dyn_plan(net, x1: torch.Tensor, confidence: str, edm) -> _DynPlan:
    prec = net.precision or rt.default_precision()
    key = ("dynamic", tuple(x1.shape), str(x1.device), prec, confidence, id(edm), bool(net.use_cuda_graph))
    p = net._plans.get(key)
    if p is None or p.generation != rt.generation():
        p = _DynPlan(net, tuple(x1.shape), x1.device, prec, confidence, edm)
        net._plans[key] = p
    return p


def run_dynamic(net, x: torch.Tensor, threshold, confidence, edm):
    if confidence not in ('edm', 'entropy', 'max'):
        raise ValueError(confidence)
    if confidence == 'edm' and (edm is False or edm is None):
        raise ValueError("confidence='edm' needs an EDM module (eval.py:110-112)")
    ys, flags, confs = [], [], []
    launches = 0
    for n in range(x.shape[0]):
        x1 = x[n:n + 1]
        plan = _get_dyn_plan(net, x1, confidence, edm)
        plan.x_static.copy_(x1)
        taken = None
        conf_val = None
        for k, _ in enumerate(plan.exits):
            plan.trunks[k].run()
            launches += plan.trunks[k].n_launches
            if confidence == 'edm':
                conf_t = plan.conf[k]
                conf_val = conf_t.view(1, 1).clone()
                if float(conf_t.item()) > threshold:      # host decision = the reference's implicit sync
                    continue
                plan.heads[k].run()
                launches += plan.heads[k].n_launches
                taken = k
                break
            else:
                plan.heads[k].run()
                launches += plan.heads[k].n_launches
                s = _confidence(plan.outs[k], threshold if confidence == 'max' else 2.0, net._num_classes)
                if confidence == 'entropy':
                    conf_val = s[0].item()
                    if conf_val < threshold:
                        taken = k
                        break
                else:
                    conf_val = s[1].item()
                    if conf_val > threshold:
                        taken = k
                        break
        if taken is None:
            plan.tail.run()
            launches += plan.tail.n_launches
            ys.append(plan.outs[-1].clone() if x.shape[0] > 1 else plan.outs[-1])
            flags.append(0)
        else:
            ys.append(plan.outs[taken].clone() if x.shape[0] > 1 else plan.outs[taken])
            flags.append(1)
        confs.append(conf_val)
    net.last_dynamic_launches = launches
    return ys, flags, confs
