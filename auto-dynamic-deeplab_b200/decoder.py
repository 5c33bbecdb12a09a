"""DeepLabv3+ decoder — B200 drop-in for `modeling/decoder.py::Decoder` (:6-29).

`_conv` keeps the reference's Sequential indices (1,2 | 4,5 | 7) for state_dict compatibility.
The 256+48 concat is a 304-channel buffer whose low-level slice is written in place by the
producer; the two 3×3 convs store post-ReLU activations; the classifier (with bias, Q8) writes fp32
logits at decoder resolution; the final ×8 bilinear is fused with its consumer (materialised NCHW
logits, or argmax + confusion matrix / entropy) in head.cu."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import runtime as rt
from .operations import AddModule, _conv_holder
from .runtime import Builder, ConvWeights, View, RELU_IN, RELU_OUT

LOW_LEVEL_C = 48
ASPP_C = 256


class Decoder(AddModule):
    def __init__(self, n_class, BatchNorm):
        super().__init__()
        eps, momentum = 1e-5, 0.1
        self.n_class = n_class
        self._conv = nn.Sequential(
            nn.ReLU(inplace=True),
            _conv_holder(ASPP_C + LOW_LEVEL_C, 256, 3, 1, 1),
            BatchNorm(256, eps=eps, momentum=momentum),
            nn.ReLU(inplace=True),
            _conv_holder(256, 256, 3, 1, 1),
            BatchNorm(256, eps=eps, momentum=momentum),
            nn.ReLU(inplace=True),
            _conv_holder(256, n_class, 1, 1, 0, bias=True))

    def _prepare(self):
        self.cw1 = ConvWeights(self._conv[1].weight, self._conv[2])
        self.cw2 = ConvWeights(self._conv[4].weight, self._conv[5])
        self.cw3 = ConvWeights(self._conv[7].weight, None, self._conv[7].bias)

    def new_cat(self, b: Builder, n: int, h: int, w: int) -> View:
        """The 304-channel concat buffer; slice [256:304] is the low-level feature."""
        return b.alloc(n, h, w, ASPP_C + LOW_LEVEL_C)

    def emit_lowres(self, b: Builder, x: View, cat: View) -> View:
        """decoder.py:24-27 + `_conv`: returns fp32 logits [N,h,w,n_class] at decoder resolution.
        `cat[..., 256:304]` must already hold the low-level feature.  Only H is compared when
        deciding whether to resize (Q8)."""
        self._ensure_prepared()
        # `_conv` starts with ReLU and is the only reader of the concat, so both producers store relu(.) (the low-level
        # slice is written with RELU_OUT by its producer) and the 3x3 needs no ReLU-on-load pass over its A tiles
        b.bilinear(x, cat.slice(0, ASPP_C), RELU_OUT, "Decoder.up" if x.h != cat.h else "Decoder.copy")
        t1 = b.scratch(cat.n, cat.h, cat.w, 256)
        b.conv(cat, t1, self.cw1, 1, 1, 1, RELU_OUT, "Decoder.conv1")
        t2 = b.scratch(cat.n, cat.h, cat.w, 256)
        b.conv(t1, t2, self.cw2, 1, 1, 1, RELU_OUT, "Decoder.conv2")
        pad_c = (self.n_class + 3) // 4 * 4
        logits = View(b.raw((cat.n, cat.h, cat.w, pad_c), torch.float32), 0, self.n_class)
        b.conv(t2, logits, self.cw3, 1, 0, 1, 0, "Decoder.classifier")
        b.release(t1)
        b.release(t2)
        return logits

    def forward(self, x: torch.Tensor, low_level: torch.Tensor, size) -> torch.Tensor:
        """decoder.py:23-29 — stand-alone call (NCHW in, NCHW fp32 logits out)."""
        self._check_eval()
        rt.require_cuda(x)
        dtype = x.dtype if x.dtype in (torch.float32, torch.bfloat16) else torch.float32
        b = Builder(x.device, dtype, record=False)
        xv = rt.as_nhwc_view(x, b, dtype)
        lv = rt.as_nhwc_view(low_level, b, dtype)
        if x.shape[2] == low_level.shape[2] and x.shape[3] != low_level.shape[3]:
            raise RuntimeError("Decoder: equal H but different W — torch.cat fails in the reference too (Q8)")
        cat = self.new_cat(b, lv.n, lv.h, lv.w)
        b.bilinear(lv, cat.slice(ASPP_C, LOW_LEVEL_C), RELU_OUT, "Decoder.low_copy")
        logits = self.emit_lowres(b, xv, cat)
        H, W = int(size[0]), int(size[1])
        out = torch.empty((xv.n, self.n_class, H, W), device=x.device, dtype=torch.float32)
        b.upsample_logits(logits, out, H, W)
        return out
