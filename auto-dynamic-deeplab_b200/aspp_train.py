"""ASPP head used by ADD — B200 drop-in for `modeling/aspp_train.py::ASPP_train` (:8-61).

Five branches write straight into channel slices of one 5·depth buffer (no torch.cat); every
branch is one launch with ReLU-on-load, folded BN and ReLU-on-store.  The image-pool branch is
GAP → 1×1 on the pooled vector → broadcast (align_corners=True from a 1×1 source is a pure
broadcast, aspp_train.py:54-55)."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import runtime as rt
from .operations import AddModule, _conv_holder
from .runtime import Builder, ConvWeights, View, RELU_IN, RELU_OUT, IN_RELUD


class ASPP_train(AddModule):
    def __init__(self, C, out, BatchNorm, depth=256, conv=nn.Conv2d, eps=1e-5, momentum=0.1, mult=1):
        super().__init__()
        self._C, self._depth, self._out = C, depth, out
        self.dils = [int(6 * mult), int(12 * mult), int(18 * mult)]
        self.aspp1 = _conv_holder(C, depth, 1)
        self.aspp2 = _conv_holder(C, depth, 3, 1, self.dils[0], self.dils[0])
        self.aspp3 = _conv_holder(C, depth, 3, 1, self.dils[1], self.dils[1])
        self.aspp4 = _conv_holder(C, depth, 3, 1, self.dils[2], self.dils[2])
        self.aspp5 = _conv_holder(C, depth, 1)
        self.conv1 = _conv_holder(depth * 5, out, 1)
        self.bn1 = BatchNorm(out, eps=eps, momentum=momentum)
        self.aspp1_bn = BatchNorm(depth, eps=eps, momentum=momentum)
        self.aspp2_bn = BatchNorm(depth, eps=eps, momentum=momentum)
        self.aspp3_bn = BatchNorm(depth, eps=eps, momentum=momentum)
        self.aspp4_bn = BatchNorm(depth, eps=eps, momentum=momentum)
        self.aspp5_bn = BatchNorm(depth, eps=eps, momentum=momentum)

    def _prepare(self):
        self.cw = [ConvWeights(getattr(self, f"aspp{k}").weight, getattr(self, f"aspp{k}_bn")) for k in range(1, 6)]
        self.cw_out = ConvWeights(self.conv1.weight, self.bn1)
        d = self._depth
        # the 1x1 over the concatenation, split: rows of the four spatial branches / rows of the image-pool branch
        dev = self.cw_out.w.device
        self.cw_out_main = ConvWeights.from_folded(self.cw_out.w_h[:, :, :4 * d, :].contiguous(), self.cw_out.bias_h, dev)
        self.w_out_pool = rt.host_to(self.cw_out.w_h[0, 0, 4 * d:, :], dev)       # [depth][out] fp32

    def out_shape(self, n, c, h, w):
        return n, self._out, h, w

    def emit(self, b: Builder, x: View, y: View, flags: int = 0) -> None:
        """aspp_train.py:34-61."""
        self._ensure_prepared()
        d = self._depth
        rin = 0 if flags & IN_RELUD else RELU_IN      # IN_RELUD: the producer already stored relu(x) (aspp_train.py:35)
        flags &= ~IN_RELUD
        cat = b.scratch(x.n, x.h, x.w, 4 * d)
        b.conv(x, cat.slice(0, d), self.cw[0], 1, 0, 1, rin | RELU_OUT, "ASPP.aspp1")
        for i, dil in enumerate(self.dils):
            b.conv(x, cat.slice(d * (i + 1), d), self.cw[i + 1], 1, dil, dil, rin | RELU_OUT, f"ASPP.aspp{i + 2}")
        # image-pool branch (aspp_train.py:49-57): GAP -> 1x1+BN+ReLU -> broadcast is constant over the image, so it
        # enters the final 1x1 as a per-image bias instead of 256 broadcast channels of the concatenation
        pooled = b.raw((x.n, self._C), torch.float32)
        b.gap(x, pooled, rin, "ASPP.gap")
        bias_n = b.raw((x.n, self._out), torch.float32)
        b.aspp_pool_bias(pooled, self.cw[4], self.w_out_pool, self.cw_out.bias, bias_n, "ASPP.pool_bias")
        b.conv(cat, y, self.cw_out_main, 1, 0, 1, flags, "ASPP.conv1", image_bias=bias_n)
        b.release(cat)
