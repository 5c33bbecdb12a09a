"""Supernet cell — drop-ins for modeling/cell_level_search.py:

* `MixedOp` (:10-29): the weighted sum of ALL eight primitives applied to the same input, sum_k w_k * op_k(x), every op
  built with affine=False and the two pools followed by their own BatchNorm(affine=False);
* `Cell` (:32-155): the search cell with down / same / up inputs, B steps of MixedOp edges per input list, concat of the
  last B states per list;
* `softmax_rows`: the alpha / beta softmax of model_net_search.py:294-310 on the device, with autograd.

Same constructors, `_ops` layouts and state_dict keys as the reference.  Two execution modes:

* training mode (`module.train()`; what search.py runs): forward AND backward through libadd_b200 — the operator
  kernels of `training.py` plus csrc/mixed.cu: the K weighted primitive outputs are combined by ONE kernel that reads the
  architecture weights on the device (`add_weighted_sum_fwd`; its backward yields dw_k = <dy, y_k> and dy_k = w_k dy), so
  gradients reach the convolution weights, the inputs and the alphas;
* eval mode: the fused inference kernels; the four parameter-free primitives of an edge (none, max_pool+BN, avg_pool+BN,
  skip_connect) and their four weights are ONE kernel that reads x once (`add_mixed_light_fwd`), the four conv primitives
  are the BN-folded tensor-core / CUDA-core kernels, combined by the same device-weighted sum.
`forward(x, weights, training=False)` is the reference's argmax path (:26-28)."""
from __future__ import annotations

import ctypes
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa: F401  (kept for API parity; not used for compute)

from . import runtime as rt
from . import training as T
from ._lib import lib, check, AddTensor
from .genotypes import PRIMITIVES
from .operations import OPS, AddModule, ReLUConvBN, FactorizedReduce, DoubleFactorizedReduce, _Pool3x3, batch_norm_forward

TP = ctypes.POINTER(AddTensor)


# ---- autograd pieces that only the supernet needs ------------------------------------------------------------------------
class _Pool(torch.autograd.Function):
    """nn.AvgPool2d(3, stride, 1, count_include_pad=False) / nn.MaxPool2d(3, stride, 1) (operations.py:9-10)."""

    @staticmethod
    def forward(ctx, x, mode, stride):
        x = T._nhwc(x)
        n, c, h, w = x.shape
        y = T._new(n, c, (h - 1) // stride + 1, (w - 1) // stride + 1, x.device)
        check(lib.add_pool3x3_fwd(ctypes.byref(T._desc(x)), ctypes.byref(T._desc(y)), mode, stride, 0, T._stream(x.device)), "pool_fwd")
        ctx.save_for_backward(x)
        ctx.cfg = (mode, stride)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        mode, stride = ctx.cfg
        dy = T._nhwc(dy)
        dx = T._new(*x.shape, x.device)
        check(lib.add_pool3x3_bwd(ctypes.byref(T._desc(x)), ctypes.byref(T._desc(dy)), ctypes.byref(T._desc(dx)), mode, stride, 0,
                                  T._stream(x.device)), "pool_bwd")
        return dx, None, None


def _desc_array(ts: List[Optional[torch.Tensor]]):
    """HOST array of K descriptor pointers (NULL = primitive skipped); keeps the descriptors alive."""
    descs = [T._desc(t) if t is not None else None for t in ts]
    arr = (TP * len(ts))(*[ctypes.pointer(d) if d is not None else None for d in descs])
    return arr, descs


class _WeightedSum(torch.autograd.Function):
    """out = sum_k w[k] * y_k with w [K] on the device (softmaxed alphas); y_k = None is a skipped primitive ('none')."""

    @staticmethod
    def forward(ctx, w, present, *ys):
        it = iter(ys)
        full = [T._nhwc(next(it)) if p else None for p in present]
        ref = next(t for t in full if t is not None)
        out = T._new(*ref.shape, ref.device)
        wf = w.detach().to(torch.float32).contiguous()
        arr, keep = _desc_array(full)
        check(lib.add_weighted_sum_fwd(arr, len(full), wf.data_ptr(), ctypes.byref(T._desc(out)), T._stream(ref.device)), "weighted_sum_fwd")
        ctx.save_for_backward(wf, *[t for t in full if t is not None])
        ctx.present = present
        return out

    @staticmethod
    def backward(ctx, dy):
        wf, *saved = ctx.saved_tensors
        present = ctx.present
        dy = T._nhwc(dy)
        it = iter(saved)
        full = [next(it) if p else None for p in present]
        dev = dy.device
        need = [ctx.needs_input_grad[2 + i] for i in range(len(saved))]
        it_need = iter(need)
        douts = [(T._new(*dy.shape, dev) if next(it_need) else None) if p else None for p in present]
        dw = torch.zeros(len(present), dtype=torch.float32, device=dev)
        ws = T._ws(lib.add_weighted_sum_workspace_bytes(*[dy.shape[i] for i in (0, 2, 3, 1)]), dev)
        ya, k1 = _desc_array(full)
        da, k2 = _desc_array(douts)
        check(lib.add_weighted_sum_bwd(ctypes.byref(T._desc(dy)), ya, da, len(present), wf.data_ptr(), dw.data_ptr(), ws.data_ptr(),
                                       ws.numel(), T._stream(dev)), "weighted_sum_bwd")
        return (dw, None) + tuple(d for d, p in zip(douts, present) if p)


class _SoftmaxRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x2 = x.detach().to(torch.float32).contiguous().view(-1, x.shape[-1])
        y = torch.empty_like(x2)
        check(lib.add_softmax_rows_fwd(x2.data_ptr(), y.data_ptr(), x2.shape[0], x2.shape[1], T._stream(x.device)), "softmax_rows_fwd")
        ctx.save_for_backward(y)
        ctx.shape = x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        g = dy.to(torch.float32).contiguous().view(y.shape)
        dx = torch.empty_like(y)
        check(lib.add_softmax_rows_bwd(y.data_ptr(), g.data_ptr(), dx.data_ptr(), y.shape[0], y.shape[1], T._stream(dy.device)), "softmax_rows_bwd")
        return dx.view(ctx.shape)


def softmax_rows(x: torch.Tensor) -> torch.Tensor:
    """F.softmax(x, dim=-1) of the architecture parameters (alphas [edges, 8]; betas slices, model_net_search.py:294-310)
    on the device, differentiable."""
    rt.require_cuda(x, "alphas")
    return _SoftmaxRows.apply(x)


class _PoolBN(nn.Sequential):
    """nn.Sequential(pool, BatchNorm(C, affine=False)) of cell_level_search.py:20-21 (keys `<i>.1.running_*`)."""

    def forward(self, x):
        return batch_norm_forward(self[1], self[0](x))


class MixedOp(nn.Module):
    def __init__(self, C, stride, BatchNorm):
        super().__init__()
        eps, momentum = 1e-5, 0.1
        self._ops = nn.ModuleList()
        self.stride = stride
        for primitive in PRIMITIVES:
            op = OPS[primitive](C, stride, BatchNorm, eps, momentum, False)
            if 'pool' in primitive:
                op = _PoolBN(op, BatchNorm(C, eps=eps, momentum=momentum, affine=False))
            self._ops.append(op)

    # ---- training mode: autograd through our kernels ---------------------------------------------------------------------
    def _forward_train(self, x: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
        ys, present = [], []
        for name, op in zip(PRIMITIVES, self._ops):
            if name == 'none':
                present.append(False)                      # w * (x * 0): contributes exact zeros and no gradient
                continue
            present.append(True)
            if 'pool' in name:
                ys.append(T.batch_norm(op[1], _Pool.apply(x, 1 if name.startswith('max') else 0, self.stride)))
            elif name == 'skip_connect':
                ys.append(x)
            elif name.startswith('sep_conv'):
                ys.append(T.sep_conv(op, x))
            else:
                ys.append(T.dil_conv(op, x))
        return _WeightedSum.apply(weights, tuple(present), *ys)

    # ---- eval mode: fused inference kernels --------------------------------------------------------------------------------
    def _forward_eval(self, x: torch.Tensor, weights: torch.Tensor) -> torch.Tensor:
        dtype = x.dtype if x.dtype in (torch.float32, torch.bfloat16) else torch.float32
        xin = x.to(dtype).contiguous(memory_format=torch.channels_last)
        n, c, h, w = xin.shape
        dev = xin.device
        wdev = weights.detach().to(device=dev, dtype=torch.float32).contiguous()
        # the four conv primitives (PRIMITIVES[4:]): BN-folded fused kernels, one output each, then ONE device-weighted sum
        convs = [self._ops[k](xin).float().contiguous(memory_format=torch.channels_last) for k in range(4, 8)]
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=dev, memory_format=torch.channels_last)
        arr, keep = _desc_array(convs)
        check(lib.add_weighted_sum_fwd(arr, 4, wdev.data_ptr() + 16, ctypes.byref(T._desc(out)), T._stream(dev)), "mixed.conv_sum")
        # none + max_pool/BN + avg_pool/BN + skip_connect and their four weights: ONE kernel, x read once, accumulated
        stats = []
        for k in (1, 2):
            bn = self._ops[k][1]
            stats += [bn.running_mean.float().contiguous(), torch.rsqrt(bn.running_var.float() + bn.eps).contiguous()]
        x32 = xin if xin.dtype == torch.float32 else xin.float()
        from ._lib import ACCUMULATE
        check(lib.add_mixed_light_fwd(ctypes.byref(T._desc(x32)), ctypes.byref(T._desc(out)), wdev.data_ptr(), stats[0].data_ptr(),
                                      stats[1].data_ptr(), stats[2].data_ptr(), stats[3].data_ptr(), ACCUMULATE, T._stream(dev)), "mixed.light")
        return out

    def forward(self, x: torch.Tensor, weights: torch.Tensor, training: bool = True) -> torch.Tensor:
        rt.require_cuda(x)
        if not training:
            return self._ops[int(torch.argmax(weights))](x)        # cell_level_search.py:27-28
        assert weights.numel() == len(self._ops), (weights.numel(), len(self._ops))
        if self.training:
            return self._forward_train(x, weights.reshape(-1))
        if self.stride != 1:
            raise NotImplementedError("eval-mode MixedOp with stride 2 (the reference Cell only builds stride 1)")
        return self._forward_eval(x, weights.reshape(-1))


def _apply_prep(m, x, training: bool):
    """ReLUConvBN / (Double)FactorizedReduce preprocessing in the cell's mode (train: autograd kernels; eval: fused)."""
    if training:
        return T._prep(m, x)
    return m(x).float()


class Cell(nn.Module):
    """reference cell_level_search.py:32-155."""

    def __init__(self, B, prev_prev_C, prev_C_down, prev_C_same, prev_C_up, C_out, BatchNorm=nn.BatchNorm2d,
                 pre_preprocess_sample_rate=1):
        super().__init__()
        if prev_C_down is not None:
            self.preprocess_down = FactorizedReduce(prev_C_down, C_out, BatchNorm=BatchNorm, affine=False)
        if prev_C_same is not None:
            self.preprocess_same = ReLUConvBN(prev_C_same, C_out, 1, 1, 0, BatchNorm=BatchNorm, affine=False)
        if prev_C_up is not None:
            self.preprocess_up = ReLUConvBN(prev_C_up, C_out, 1, 1, 0, BatchNorm=BatchNorm, affine=False)
        if prev_prev_C != -1:
            if pre_preprocess_sample_rate >= 1:
                self.pre_preprocess = ReLUConvBN(prev_prev_C, C_out, 1, 1, 0, BatchNorm=BatchNorm, affine=False)
            elif pre_preprocess_sample_rate == 0.5:
                self.pre_preprocess = FactorizedReduce(prev_prev_C, C_out, BatchNorm=BatchNorm, affine=False)
            elif pre_preprocess_sample_rate == 0.25:
                self.pre_preprocess = DoubleFactorizedReduce(prev_prev_C, C_out, BatchNorm=BatchNorm, affine=False)
        self.B = B
        self._ops = nn.ModuleList()
        for i in range(self.B):
            for j in range(2 + i):
                if prev_prev_C == -1 and j == 0:
                    op = None
                else:
                    op = MixedOp(C_out, 1, BatchNorm)
                self._ops.append(op)

    def scale_dimension(self, dim, scale):
        assert isinstance(dim, int)
        return int((float(dim) - 1.0) * scale + 1.0) if dim % 2 else int(dim * scale)

    def prev_feature_resize(self, prev_feature, mode):
        s = 0.5 if mode == 'down' else 2
        size = (self.scale_dimension(prev_feature.shape[2], s), self.scale_dimension(prev_feature.shape[3], s))
        return self._bilinear(prev_feature, size)

    def _bilinear(self, x, size):
        if self.training:
            return T.bilinear(x, size)
        with torch.no_grad():
            return T.bilinear(x.float(), size)                    # the forward kernel alone (no tape)

    def forward(self, s0, s1_down, s1_same, s1_up, n_alphas):
        rt.require_cuda(n_alphas, "n_alphas")
        tr = self.training
        size_h = size_w = None
        if s1_down is not None:
            s1_down = _apply_prep(self.preprocess_down, s1_down, tr)
            size_h, size_w = s1_down.shape[2], s1_down.shape[3]
        if s1_same is not None:
            s1_same = _apply_prep(self.preprocess_same, s1_same, tr)
            size_h, size_w = s1_same.shape[2], s1_same.shape[3]
        if s1_up is not None:
            s1_up = self.prev_feature_resize(s1_up, 'up')
            s1_up = _apply_prep(self.preprocess_up, s1_up, tr)
            size_h, size_w = s1_up.shape[2], s1_up.shape[3]
        all_states = []
        if s0 is not None:
            if s0.shape[2] < size_h or s0.shape[3] < size_w:
                s0 = self._bilinear(s0, (size_h, size_w))
            s0 = _apply_prep(self.pre_preprocess, s0, tr)
            first = s0
        else:
            first = 0
        for s1 in (s1_down, s1_same, s1_up):
            if s1 is not None:
                all_states.append([first, s1])
        final_concates = []
        for states in all_states:
            offset = 0
            for i in range(self.B):
                new_states = []
                for j, h in enumerate(states):
                    branch_index = offset + j
                    if self._ops[branch_index] is None:
                        continue
                    new_states.append(self._ops[branch_index](h, n_alphas[branch_index]))
                s = new_states[0]
                for t in new_states[1:]:
                    s = T.add(s, t) if tr else _add_eval(s, t)
                offset += len(states)
                states.append(s)
            last = states[-self.B:]
            final_concates.append(T.cat(last) if tr else _cat_eval(last))
        return final_concates


def _add_eval(a, b):
    with torch.no_grad():
        return T.add(a, b)


def _cat_eval(xs):
    with torch.no_grad():
        return T.cat(xs)
