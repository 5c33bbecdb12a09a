"""MixedOp — drop-in for modeling/cell_level_search.py:10-29 (the supernet edge): the weighted sum
of ALL eight primitives applied to the same input,  sum_k w_k * op_k(x),  each op built with affine=False and the two
pools followed by their own BatchNorm(affine=False).  Same constructor, `_ops` layout and state_dict keys as the
reference.  Every primitive runs on libadd_b200 (training mode: raw-weight conv kernels + batch-statistics BatchNorm
kernels; eval mode: the fused BN-folded kernels) and the weighted sum is accumulated on the device with
`add_scale_fwd` (y += w_k * op_k(x)).  Forward only (no autograd); the alpha / beta softmax that produces `weights`
(model_net_search.py:294-310) is a 8-element host-side torch.softmax and stays with the caller."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import runtime as rt
from .genotypes import PRIMITIVES
from .operations import OPS, AddModule, _Pool3x3, batch_norm_forward
from .runtime import ACCUMULATE, Builder


class _PoolBN(nn.Sequential):
    """nn.Sequential(pool, BatchNorm(C, affine=False)) of cell_level_search.py:20-21 (keys `<i>.1.running_*`)."""

    def forward(self, x):
        return batch_norm_forward(self[1], self[0](x))


class MixedOp(nn.Module):
    def __init__(self, C, stride, BatchNorm):
        super().__init__()
        eps, momentum = 1e-5, 0.1
        self._ops = nn.ModuleList()
        for primitive in PRIMITIVES:
            op = OPS[primitive](C, stride, BatchNorm, eps, momentum, False)
            if 'pool' in primitive:
                op = _PoolBN(op, BatchNorm(C, eps=eps, momentum=momentum, affine=False))
            self._ops.append(op)

    def forward(self, x: torch.Tensor, weights: torch.Tensor, training: bool = True) -> torch.Tensor:
        rt.require_cuda(x)
        if not training:
            return self._ops[int(torch.argmax(weights))](x)        # cell_level_search.py:27-28
        w = [float(v) for v in weights.detach().float().reshape(-1).tolist()]
        assert len(w) == len(self._ops), (len(w), len(self._ops))
        dtype = x.dtype if x.dtype in (torch.float32, torch.bfloat16) else torch.float32
        b = Builder(x.device, dtype, record=False)
        out = None
        for k, (wk, op) in enumerate(zip(w, self._ops)):
            y = op(x)                                              # each primitive: its own fused / training kernels
            yv = rt.as_nhwc_view(y, b, dtype)
            if out is None:
                out = b.alloc(yv.n, yv.h, yv.w, yv.c)
            b.scale(yv, out, wk, 1, ACCUMULATE if k > 0 else 0, f"MixedOp.w{k}")
        return out.nchw()
