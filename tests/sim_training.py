"""TEST INFRASTRUCTURE ONLY — plain-PyTorch (CPU, autograd) stand-ins for the PRIMITIVES of `add_b200.training`
(conv2d, depthwise, batch_norm, bilinear, cat, add, global average pool, broadcast, cross entropy), for checking the wiring
ABOVE them — `relu_conv_bn` … `cell_forward`, `aspp_forward`, `decoder_forward`, `add_forward`, `add_loss`, and
`ADD.forward` in `.train()` — against the train-step fixtures of the unmodified reference without a GPU.  The product never
imports this file; the kernels and their backward are pinned by tests/test_gpu_training.py on the B200."""
from __future__ import annotations

import torch
import torch.nn.functional as F

from add_b200 import training as T
from add_b200 import runtime as rt


def _pad_or_crop(x, pad):
    return F.pad(x, (pad, pad, pad, pad)) if pad >= 0 else x[:, :, -pad:, -pad:]


def conv2d(x, weight, bias=None, stride=1, pad=0, dil=1, relu_in=False, cout_pad=None):
    xin = torch.relu(x) if relu_in else x
    if xin.shape[1] > weight.shape[1]:          # the image is carried with a zero 4th channel (16-byte pixels): no weights for it
        xin = xin[:, :weight.shape[1]]
    if pad >= 0:
        y = F.conv2d(xin, weight, bias, stride, pad, dil)
    else:
        # FactorizedReduce's odd lattice: pad(x, (0,1,0,1))[:, :, 1:, 1:] then a stride-s 1x1 conv (operations.py:97-99)
        xs = F.pad(xin, (0, -pad, 0, -pad))[:, :, -pad:, -pad:]
        y = F.conv2d(xs, weight, bias, stride, 0, dil)
    cout = weight.shape[0]
    if cout_pad is not None and cout_pad > cout:
        y = F.pad(y, (0, 0, 0, 0, 0, cout_pad - cout))
    return y


def depthwise(x, weight, relu_in=False):
    xin = torch.relu(x) if relu_in else x
    return F.conv2d(xin, weight, None, 1, weight.shape[2] // 2, 1, x.shape[1])


def batch_norm(bn, x, relu=False, sync=None, group=None):
    w, b = (bn.weight, bn.bias) if bn.affine else (None, None)
    y = F.batch_norm(x, bn.running_mean, bn.running_var, w, b, True, bn.momentum, bn.eps)
    if bn.track_running_stats and bn.num_batches_tracked is not None:
        bn.num_batches_tracked += 1
    return torch.relu(y) if relu else y


def bilinear(x, size):
    if x.shape[2] == size[0] and x.shape[3] == size[1]:
        return x
    return F.interpolate(x, (int(size[0]), int(size[1])), mode="bilinear", align_corners=False)


class _GlobalAvgPool:
    @staticmethod
    def apply(x, relu_in):
        return (torch.relu(x) if relu_in else x).mean(dim=(2, 3), keepdim=True)


class _Broadcast:
    @staticmethod
    def apply(x, ho, wo):
        return x.expand(x.shape[0], x.shape[1], ho, wo)


def cross_entropy(logits, target, num_class=19, ignore_index=255, class_weight=None):
    return F.cross_entropy(logits[:, :num_class], target.long(), class_weight, ignore_index=ignore_index)


def install(monkeypatch) -> None:
    for name, obj in dict(conv2d=conv2d, depthwise=depthwise, batch_norm=batch_norm, bilinear=bilinear,
                          cat=lambda xs: torch.cat(list(xs), 1), add=lambda a, b: a + b, _GlobalAvgPool=_GlobalAvgPool,
                          _Broadcast=_Broadcast, cross_entropy=cross_entropy).items():
        monkeypatch.setattr(T, name, obj)
    monkeypatch.setattr(rt, "require_cuda", lambda *a, **k: None)
