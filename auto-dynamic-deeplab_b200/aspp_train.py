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
from .runtime import Builder, ConvWeights, View, RELU_IN, RELU_OUT


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

    def out_shape(self, n, c, h, w):
        return n, self._out, h, w

    def emit(self, b: Builder, x: View, y: View, flags: int = 0) -> None:
        """aspp_train.py:34-61."""
        self._ensure_prepared()
        d = self._depth
        cat = b.scratch(x.n, x.h, x.w, 5 * d)
        b.conv(x, cat.slice(0, d), self.cw[0], 1, 0, 1, RELU_IN | RELU_OUT, "ASPP.aspp1")
        for i, dil in enumerate(self.dils):
            b.conv(x, cat.slice(d * (i + 1), d), self.cw[i + 1], 1, dil, dil, RELU_IN | RELU_OUT, f"ASPP.aspp{i + 2}")
        pooled = View(b.raw((x.n, 1, 1, self._C), torch.float32))
        b.gap(x, pooled.buf, RELU_IN, "ASPP.gap")
        p5 = View(b.raw((x.n, 1, 1, d), torch.float32))
        b.conv(pooled, p5, self.cw[4], 1, 0, 1, RELU_OUT, "ASPP.aspp5")
        b.bilinear(p5, cat.slice(4 * d, d), 0, "ASPP.broadcast")
        b.conv(cat, y, self.cw_out, 1, 0, 1, flags, "ASPP.conv1")
        b.release(cat)
