"""Operator zoo — B200 drop-ins for the reference's `modeling/operations.py`.

Same public names, constructor signatures and `state_dict` keys as the reference
(operations.py:7-180) so `load_state_dict(reference_checkpoint)` works unchanged; the compute is
different: every module *emits* fused libadd_b200 launches (ReLU-on-load, BN folded into the conv
weights, node-sum as accumulate-into-slice) instead of calling ATen ops.  `torch.nn.Conv2d` /
`BatchNorm2d` objects appear here only as parameter containers that fix the state_dict layout;
their `forward` is never invoked.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

from . import runtime as rt
from .runtime import Builder, ConvWeights, View, RELU_IN, RELU_OUT, ACCUMULATE, IN_RELUD
from ._lib import lib, check


from .sync_batchnorm import SynchronizedBatchNorm2d, batch_norm_forward  # noqa: E402  (reference name, modeling/sync_batchnorm/batchnorm.py:180)


class AddModule(nn.Module):
    """Base: generation-tracked weight cache + the emit/forward protocol."""

    STATELESS = False          # True: no parameters / statistics, so .train() and .eval() compute the same thing

    def __init__(self):
        super().__init__()
        self._prep_gen = -1

    # any parameter movement / reload invalidates folded weights and recorded plans
    def _apply(self, fn, *a, **k):
        rt.bump_generation()
        return super()._apply(fn, *a, **k)

    def _load_from_state_dict(self, *a, **k):
        rt.bump_generation()
        return super()._load_from_state_dict(*a, **k)

    def invalidate(self) -> None:
        """Call after mutating parameters in place (e.g. an optimizer step)."""
        rt.bump_generation()

    def train(self, mode: bool = True):
        # leaving training mode is where in-place parameter updates (any optimiser, torch's included) and moved running
        # statistics meet the fused inference path: whatever was folded / recorded before is stale from here on
        if bool(mode) != self.training:
            rt.bump_generation()
        return super().train(mode)

    def _ensure_prepared(self) -> None:
        if self._prep_gen != rt.generation():
            self._prepare()
            self._prep_gen = rt.generation()

    def _prepare(self) -> None:  # fold BN, permute weights
        raise NotImplementedError

    def _check_eval(self) -> None:
        if self.training:
            raise NotImplementedError(
                f"{type(self).__name__}: training-mode forward (batch statistics / autograd) is not part of "
                "the accelerated inference path; call .eval() (SURVEY §8f lists training as 'next')")

    # ---- stand-alone call: NCHW in → NCHW (channels_last) out -------------------------------
    def out_shape(self, n, c, h, w):
        return n, c, h, w

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.training and not self.STATELESS:
            y = self._forward_train(x)
            rt.bump_generation()          # running statistics moved: folded eval-mode weights / recorded plans are stale
            return y
        rt.require_cuda(x)
        dtype = x.dtype if x.dtype in (torch.float32, torch.bfloat16) else torch.float32
        b = Builder(x.device, dtype, record=False)
        xv = rt.as_nhwc_view(x, b, dtype)
        n, c, h, w = self.out_shape(*x.shape)
        yv = b.alloc(n, h, w, c)
        self.emit(b, xv, yv, 0)
        return yv.nchw()

    def emit(self, b: Builder, x: View, y: View, flags: int = 0) -> None:
        raise NotImplementedError

    # ---- training-mode forward: batch statistics, unfused, eager ------------------------
    # The conv kernels run with the RAW weights (no BN fold) and every BatchNorm is the three-kernel batch-statistics
    # forward of sync_batchnorm.batch_norm_forward (synchronised over the process group when there is one).  Forward
    # only — no autograd graph is built.  Modules that have no training forward yet raise.
    def _forward_train(self, x: torch.Tensor) -> torch.Tensor:
        self._check_eval()

    def _train_io(self, x: torch.Tensor):
        rt.require_cuda(x)
        dtype = x.dtype if x.dtype in (torch.float32, torch.bfloat16) else torch.float32
        b = Builder(x.device, dtype, record=False)
        return b, rt.as_nhwc_view(x, b, dtype)


def _relu_in(flags: int) -> int:
    """ReLU-on-load unless the producer already stored relu(x) (IN_RELUD); strips the host-only bit."""
    return (0 if flags & IN_RELUD else RELU_IN) | (flags & ~IN_RELUD)


def _conv_holder(cin, cout, k, stride=1, padding=0, dilation=1, groups=1, bias=False) -> nn.Conv2d:
    return nn.Conv2d(cin, cout, k, stride=stride, padding=padding, dilation=dilation, groups=groups, bias=bias)


class ReLUConvBN(AddModule):
    """ReLU → conv → BN as ONE launch (reference: operations.py:18-29; keys op.1.weight, op.2.*)."""

    def __init__(self, C_in, C_out, kernel_size, stride, padding, BatchNorm, eps=1e-5, momentum=0.1, affine=True):
        super().__init__()
        self.op = nn.Sequential(nn.ReLU(inplace=False),
                                _conv_holder(C_in, C_out, kernel_size, stride, padding),
                                BatchNorm(C_out, eps=eps, momentum=momentum, affine=affine))
        self.stride, self.padding, self.C_out = stride, padding, C_out

    def _prepare(self):
        self.cw = ConvWeights(self.op[1].weight, self.op[2])

    def out_shape(self, n, c, h, w):
        k = self.op[1].kernel_size[0]
        return (n, self.C_out, (h + 2 * self.padding - k) // self.stride + 1,
                (w + 2 * self.padding - k) // self.stride + 1)

    def emit(self, b, x, y, flags=0):
        self._ensure_prepared()
        b.conv(x, y, self.cw, self.stride, self.padding, 1, _relu_in(flags), "ReLUConvBN")

    def _forward_train(self, x):
        b, xv = self._train_io(x)
        n, c, h, w = self.out_shape(*x.shape)
        t = b.alloc(n, h, w, c)
        b.conv(xv, t, ConvWeights(self.op[1].weight), self.stride, self.padding, 1, RELU_IN, "ReLUConvBN.train")
        return batch_norm_forward(self.op[2], t.nchw())


class DilConv(AddModule):
    """ReLU → DENSE C→C k×k dilated conv → BN as one implicit-GEMM launch
    (reference: operations.py:32-43 — no `groups`, so this is NOT depthwise)."""

    def __init__(self, C_in, C_out, kernel_size, stride, padding, dilation, BatchNorm, eps=1e-5, momentum=0.1, affine=True):
        super().__init__()
        self.op = nn.Sequential(nn.ReLU(inplace=False),
                                _conv_holder(C_in, C_out, kernel_size, stride, padding, dilation),
                                BatchNorm(C_out, eps=eps, momentum=momentum, affine=affine))
        self.stride, self.padding, self.dilation, self.C_out, self.k = stride, padding, dilation, C_out, kernel_size

    def _prepare(self):
        self.cw = ConvWeights(self.op[1].weight, self.op[2])

    def out_shape(self, n, c, h, w):
        e = self.dilation * (self.k - 1) + 1
        return (n, self.C_out, (h + 2 * self.padding - e) // self.stride + 1, (w + 2 * self.padding - e) // self.stride + 1)

    def emit(self, b, x, y, flags=0):
        self._ensure_prepared()
        b.conv(x, y, self.cw, self.stride, self.padding, self.dilation, _relu_in(flags), "DilConv")

    def _forward_train(self, x):
        b, xv = self._train_io(x)
        n, c, h, w = self.out_shape(*x.shape)
        t = b.alloc(n, h, w, c)
        b.conv(xv, t, ConvWeights(self.op[1].weight), self.stride, self.padding, self.dilation, RELU_IN, "DilConv.train")
        return batch_norm_forward(self.op[2], t.nchw())


class SepConv(AddModule):
    """(ReLU → depthwise k×k → pointwise 1×1 → BN) ×2 as TWO fused launches; the mid activation is
    stored post-ReLU (reference: operations.py:46-62; keys op.1/2/3/5/6/7)."""

    def __init__(self, C_in, C_out, kernel_size, stride, padding, BatchNorm, eps=1e-5, momentum=0.1, affine=True):
        super().__init__()
        if stride != 1 or C_in != C_out:
            raise NotImplementedError("SepConv: ADD only instantiates stride 1, C_in == C_out (ADD.py:61)")
        self.op = nn.Sequential(
            nn.ReLU(inplace=False),
            _conv_holder(C_in, C_in, kernel_size, stride, padding, groups=C_in),
            _conv_holder(C_in, C_out, 1),
            BatchNorm(C_out, eps=eps, momentum=momentum, affine=affine),
            nn.ReLU(inplace=False),
            _conv_holder(C_out, C_out, kernel_size, 1, padding, groups=C_out),
            _conv_holder(C_out, C_out, 1),
            BatchNorm(C_out, eps=eps, momentum=momentum, affine=affine))
        self.k, self.C = kernel_size, C_out

    def _prepare(self):
        def dw(conv):  # [C,1,k,k] -> [k][k][C], permuted on the host
            return rt.host_to(conv.weight.detach().cpu().float()[:, 0].permute(1, 2, 0), conv.weight.device)
        self.dw1, self.dw2 = dw(self.op[1]), dw(self.op[5])
        self.pw1 = ConvWeights(self.op[2].weight, self.op[3])
        self.pw2 = ConvWeights(self.op[6].weight, self.op[7])

    def emit(self, b, x, y, flags=0):
        self._ensure_prepared()
        mid = b.scratch(x.n, x.h, x.w, self.C)
        b.sepconv_half(x, mid, self.dw1, self.pw1, self.k, _relu_in(flags & IN_RELUD) | RELU_OUT, "SepConv.half1")
        b.sepconv_half(mid, y, self.dw2, self.pw2, self.k, flags & ~IN_RELUD, "SepConv.half2")
        b.release(mid)

    def _forward_train(self, x):
        def dw(conv):  # [C,1,k,k] -> [k][k][C], permuted on the host
            return rt.host_to(conv.weight.detach().cpu().float()[:, 0].permute(1, 2, 0), conv.weight.device)
        b, xv = self._train_io(x)
        t1 = b.alloc(xv.n, xv.h, xv.w, self.C)
        b.sepconv_half(xv, t1, dw(self.op[1]), ConvWeights(self.op[2].weight), self.k, RELU_IN, "SepConv.half1.train")
        m = batch_norm_forward(self.op[3], t1.nchw(), relu=True)            # BN -> ReLU (op.4) fused in the normalise kernel
        t2 = b.alloc(xv.n, xv.h, xv.w, self.C)
        b.sepconv_half(rt.as_nhwc_view(m, b, xv.dtype), t2, dw(self.op[5]), ConvWeights(self.op[6].weight), self.k, 0,
                       "SepConv.half2.train")
        return batch_norm_forward(self.op[7], t2.nchw())


class Identity(AddModule):
    """operations.py:65-71 (skip_connect).  Stand-alone it returns its input (like the reference); as a cell edge it
    is one copy / accumulate launch."""

    def _prepare(self):
        pass

    def forward(self, x):
        return x

    def emit(self, b, x, y, flags=0):
        b.scale(x, y, 1.0, 1, flags & ~IN_RELUD, "Identity")


class Zero(AddModule):
    """operations.py:74-83 ('none'): x[:, :, ::stride, ::stride].mul(0.) — IEEE x*0, so NaN/inf propagate as in torch."""
    STATELESS = True

    def __init__(self, stride):
        super().__init__()
        self.stride = stride

    def _prepare(self):
        pass

    def out_shape(self, n, c, h, w):
        return n, c, (h - 1) // self.stride + 1, (w - 1) // self.stride + 1

    def emit(self, b, x, y, flags=0):
        b.scale(x, y, 0.0, self.stride, flags & ~IN_RELUD, "Zero")


class _Pool3x3(AddModule):
    """nn.AvgPool2d(3, stride, padding=1, count_include_pad=False) / nn.MaxPool2d(3, stride, padding=1)
    (operations.py:9-10) — parameter-free, one launch."""
    MODE = 0
    STATELESS = True

    def __init__(self, stride):
        super().__init__()
        self.stride = stride

    def _prepare(self):
        pass

    def out_shape(self, n, c, h, w):
        return n, c, (h - 1) // self.stride + 1, (w - 1) // self.stride + 1

    def emit(self, b, x, y, flags=0):
        b.pool3x3(x, y, self.MODE, self.stride, flags & ~IN_RELUD, type(self).__name__)


class AvgPool3x3(_Pool3x3):
    MODE = 0


class MaxPool3x3(_Pool3x3):
    MODE = 1


# operations.py:7-16 — same keys, same lambda signature.
OPS = {
    'none': lambda C, stride, BatchNorm, eps, momentum, affine: Zero(stride),
    'avg_pool_3x3': lambda C, stride, BatchNorm, eps, momentum, affine: AvgPool3x3(stride),
    'max_pool_3x3': lambda C, stride, BatchNorm, eps, momentum, affine: MaxPool3x3(stride),
    'skip_connect': lambda C, stride, BatchNorm, eps, momentum, affine: Identity(),
    'sep_conv_3x3': lambda C, stride, BatchNorm, eps, momentum, affine: SepConv(C, C, 3, stride, 1, BatchNorm, eps=eps, momentum=momentum, affine=affine),
    'sep_conv_5x5': lambda C, stride, BatchNorm, eps, momentum, affine: SepConv(C, C, 5, stride, 2, BatchNorm, eps=eps, momentum=momentum, affine=affine),
    'dil_conv_3x3': lambda C, stride, BatchNorm, eps, momentum, affine: DilConv(C, C, 3, stride, 2, 2, BatchNorm, eps=eps, momentum=momentum, affine=affine),
    'dil_conv_5x5': lambda C, stride, BatchNorm, eps, momentum, affine: DilConv(C, C, 5, stride, 4, 2, BatchNorm, eps=eps, momentum=momentum, affine=affine),
}


class _FactorizedReduceBase(AddModule):
    """ReLU; two stride-s 1×1 convs on the lattices at offset 0 and s/2; channel-cat; BN — as two
    launches that write the two channel halves directly, each with its half of the BN folded in.
    The odd lattice is the same kernel with pad = -s/2 (out-of-range taps read zero, matching the
    reference's ConstantPad2d + slice)."""
    STEP = 2

    def __init__(self, C_in, C_out, BatchNorm, eps=1e-5, momentum=0.1, affine=True, _bn_kwargs=None):
        super().__init__()
        assert C_out % 2 == 0
        self.relu = nn.ReLU(inplace=False)
        self.conv_1 = _conv_holder(C_in, C_out // 2, 1, stride=self.STEP)
        self.conv_2 = _conv_holder(C_in, C_out // 2, 1, stride=self.STEP)
        self.bn = BatchNorm(C_out, **(_bn_kwargs if _bn_kwargs is not None else dict(eps=eps, momentum=momentum, affine=affine)))
        self.C_out = C_out

    def _prepare(self):
        half = self.C_out // 2

        class _Half:  # a BN view restricted to one half of the channels
            def __init__(s, lo, hi, bn):
                s.running_var, s.running_mean = bn.running_var[lo:hi], bn.running_mean[lo:hi]
                s.weight = bn.weight[lo:hi] if bn.weight is not None else None
                s.bias = bn.bias[lo:hi] if bn.bias is not None else None
                s.eps = bn.eps
        self.cw1 = ConvWeights(self.conv_1.weight, _Half(0, half, self.bn))
        self.cw2 = ConvWeights(self.conv_2.weight, _Half(half, self.C_out, self.bn))
        # FactorizedReduce (stride 2) as ONE 2x2 stride-2 conv: tap (0,0) carries conv_1 into the first half of the
        # output channels, tap (1,1) carries conv_2 (the odd lattice x[2i+1, 2j+1], zero outside the image like the
        # reference's pad + slice) into the second half; taps (0,1),(1,0) are zero.  One launch, aligned channel
        # range (the halves alone start at C_out/2, which is not 16-byte aligned for C_out = 40).
        self.cw_merged = None
        if self.STEP == 2:
            cin = self.cw1.cin
            w = torch.zeros(2, 2, cin, self.C_out, dtype=torch.float32)
            w[0, 0, :, :half] = self.cw1.w_h[0, 0]
            w[1, 1, :, half:] = self.cw2.w_h[0, 0]
            self.cw_merged = ConvWeights.from_folded(w, torch.cat([self.cw1.bias_h, self.cw2.bias_h]).contiguous(),
                                                     self.cw1.w.device)

    def out_shape(self, n, c, h, w):
        s = self.STEP
        return n, self.C_out, (h - 1) // s + 1, (w - 1) // s + 1

    def emit(self, b, x, y, flags=0):
        self._ensure_prepared()
        half = self.C_out // 2
        if self.cw_merged is not None and x.dtype == torch.bfloat16 and rt.tc_available():
            b.conv(x, y, self.cw_merged, 2, 0, 1, _relu_in(flags), type(self).__name__ + ".merged")
            return
        b.conv(x, y.slice(0, half), self.cw1, self.STEP, 0, 1, _relu_in(flags), type(self).__name__ + ".even")
        b.conv(x, y.slice(half, half), self.cw2, self.STEP, -(self.STEP // 2), 1, _relu_in(flags),
               type(self).__name__ + ".odd")

    def _forward_train(self, x):
        b, xv = self._train_io(x)
        n, c, h, w = self.out_shape(*x.shape)
        half = self.C_out // 2
        t = b.alloc(n, h, w, c)
        b.conv(xv, t.slice(0, half), ConvWeights(self.conv_1.weight), self.STEP, 0, 1, RELU_IN, type(self).__name__ + ".even.train")
        b.conv(xv, t.slice(half, half), ConvWeights(self.conv_2.weight), self.STEP, -(self.STEP // 2), 1, RELU_IN,
               type(self).__name__ + ".odd.train")
        return batch_norm_forward(self.bn, t.nchw())


class FactorizedReduce(_FactorizedReduceBase):
    """operations.py:86-101."""
    STEP = 2


class DoubleFactorizedReduce(_FactorizedReduceBase):
    """operations.py:104-119 — stride 4, offset 2; its BN ignores the eps/momentum args (Q10)."""
    STEP = 4

    def __init__(self, C_in, C_out, BatchNorm, eps=1e-5, momentum=0.1, affine=True):
        super().__init__(C_in, C_out, BatchNorm, _bn_kwargs=dict(affine=affine))


class ASPP(AddModule):
    """The search-time head (reference operations.py:122-158): relu; conv11 (1x1+BN+ReLU), conv33 (dilated 3x3+BN+ReLU),
    conv_p on the global-average-pooled map (1x1+BN+ReLU, up-sampled with align_corners=True = broadcast); cat; concate_conv
    (1x1+BN+ReLU); final_conv (1x1, no BN, no bias).  Same attribute names / state_dict keys.  Training mode runs (and
    differentiates) through `training.py`; eval mode is six fused launches (the pooled branch enters the concat 1x1 as a
    per-image bias, like ASPP_train)."""

    def __init__(self, in_channels, out_channels, paddings, dilations, BatchNorm=nn.BatchNorm2d, momentum=0.0003):
        super().__init__()
        C = in_channels
        self.relu = nn.ReLU()
        self.conv11 = nn.Sequential(_conv_holder(C, C, 1), BatchNorm(C), nn.ReLU(inplace=True))
        self.conv33 = nn.Sequential(_conv_holder(C, C, 3, padding=paddings, dilation=dilations), BatchNorm(C), nn.ReLU(inplace=True))
        self.conv_p = nn.Sequential(_conv_holder(C, C, 1), BatchNorm(C), nn.ReLU(inplace=True))
        self.concate_conv = nn.Sequential(_conv_holder(C * 3, C, 1), BatchNorm(C), nn.ReLU(inplace=True))
        self.final_conv = _conv_holder(C, out_channels, 1)
        self._C, self._out, self._pad, self._dil = C, out_channels, paddings, dilations

    def _prepare(self):
        self.cw11 = ConvWeights(self.conv11[0].weight, self.conv11[1])
        self.cw33 = ConvWeights(self.conv33[0].weight, self.conv33[1])
        self.cwp = ConvWeights(self.conv_p[0].weight, self.conv_p[1])
        cwc = ConvWeights(self.concate_conv[0].weight, self.concate_conv[1])
        C, dev = self._C, cwc.w.device
        self.cwc_main = ConvWeights.from_folded(cwc.w_h[:, :, :2 * C, :].contiguous(), cwc.bias_h, dev)
        self.wc_pool = rt.host_to(cwc.w_h[0, 0, 2 * C:, :], dev)
        self.bias_c = cwc.bias
        self.cwf = ConvWeights(self.final_conv.weight)

    def out_shape(self, n, c, h, w):
        return n, self._out, h, w

    def emit(self, b, x, y, flags=0):
        self._ensure_prepared()
        C = self._C
        rin = 0 if flags & IN_RELUD else RELU_IN
        cat = b.scratch(x.n, x.h, x.w, 2 * C)
        b.conv(x, cat.slice(0, C), self.cw11, 1, 0, 1, rin | RELU_OUT, "ASPP.conv11")
        b.conv(x, cat.slice(C, C), self.cw33, 1, self._pad, self._dil, rin | RELU_OUT, "ASPP.conv33")
        pooled = b.raw((x.n, C), torch.float32)
        b.gap(x, pooled, rin, "ASPP.gap")
        bias_n = b.raw((x.n, C), torch.float32)
        b.aspp_pool_bias(pooled, self.cwp, self.wc_pool, self.bias_c, bias_n, "ASPP.pool_bias")
        t = b.scratch(x.n, x.h, x.w, C)
        b.conv(cat, t, self.cwc_main, 1, 0, 1, RELU_OUT, "ASPP.concate_conv", image_bias=bias_n)
        b.conv(t, y, self.cwf, 1, 0, 1, 0, "ASPP.final_conv")
        b.release(cat)
        b.release(t)

    def _forward_train(self, x):
        from . import training as T
        h, w = x.shape[2], x.shape[3]
        a = T.batch_norm(self.conv11[1], T.conv2d(x, self.conv11[0].weight, relu_in=True), relu=True)
        c33 = self.conv33[0]
        bq = T.batch_norm(self.conv33[1], T.conv2d(x, c33.weight, None, 1, c33.padding[0], c33.dilation[0], relu_in=True), relu=True)
        p = T._GlobalAvgPool.apply(x, True)
        p = T.batch_norm(self.conv_p[1], T.conv2d(p, self.conv_p[0].weight), relu=True)
        cat = T.cat([a, bq, T._Broadcast.apply(p, h, w)])
        y = T.batch_norm(self.concate_conv[1], T.conv2d(cat, self.concate_conv[0].weight), relu=True)
        return T.conv2d(y, self.final_conv.weight)


# ---- confidence scalars (operations.py:161-180) ------------------------------------------------

def _confidence(x: torch.Tensor, threshold: float, num_class: int):
    rt.require_cuda(x, "logits")
    n, c, h, w = x.shape
    assert c == num_class
    x = x.float().contiguous()
    nbytes = lib.add_confidence_workspace_bytes(n, h, w)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    out = torch.empty(2, dtype=torch.float32, device=x.device)
    import ctypes
    s = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    check(lib.add_confidence_nchw(x.data_ptr(), n, c, h, w, float(threshold), out.data_ptr(), ws.data_ptr(),
                                  nbytes, s), "confidence_nchw")
    return out


def normalized_shannon_entropy(x, num_class=19):
    """operations.py:161-170: -Σ p·log p / log(num_class), summed over batch and pixels, / (H·W);
    returned as a Python float (a device sync, as in the reference's `.item()`)."""
    return _confidence(x, 2.0, num_class)[0].item()


def confidence_max(x, thresold, num_class=19):
    """operations.py:172-180: fraction of pixels whose max softmax probability exceeds `thresold`."""
    return _confidence(x, thresold, num_class)[1].item()
