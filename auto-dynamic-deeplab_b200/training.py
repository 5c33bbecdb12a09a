"""Training step of the fixed-architecture ADD network (reference train.py:216-247):

    outputs = model(image)                      # ADD.forward in train mode: batch-statistics BatchNorm
    loss = mean_k CE(outputs[k], target)        # ignore_index 255, optional class weights (train.py:229-233, utils/loss.py)
    loss.backward()
    optimizer.step()                            # SGD momentum 0.9, weight decay, nesterov (train.py:126-127); poly LR

Everything that touches an activation or a gradient map is a libadd_b200 kernel (forward: the fp32 conv / depthwise /
bilinear / BatchNorm kernels of the inference library with raw weights; backward: csrc/backward.cu).  The tape is
torch.autograd: each primitive below is an `autograd.Function` whose forward and backward launch our kernels; autograd
itself only sums the gradients of tensors with several consumers and hands parameter gradients to `.grad`.  fp32 NHWC
(channels_last) throughout — this is the parity path against the reference's `.backward()`; all reductions are
fixed-order, so a step is bit-reproducible.  SynchronizedBatchNorm2d statistics (forward [sum x, sum x^2, n], backward
[sum dy, sum dy*xhat]) are exchanged through `sync_batchnorm.exchange_sum` — one fused exchange per BN layer and
direction (NVLink peer memory on GPUs, `dist.all_reduce` on gloo).

Mirrors, in train mode, reference modeling/operations.py (ReLUConvBN :18-29, DilConv :32-43, SepConv :46-62,
FactorizedReduce :86-101, DoubleFactorizedReduce :104-119), ADD.py (Cell.forward :69-116, ADD.forward :277-325),
aspp_train.py:34-61, decoder.py:23-29."""
from __future__ import annotations

import ctypes
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from ._lib import lib, check, AddTensor, ADD_F32, RELU_IN, RELU_OUT, ACCUMULATE
from . import sync_batchnorm as sbn

CL = torch.channels_last


# ---- plumbing --------------------------------------------------------------------------------------------------------
def _stream(dev) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _pix_stride(t: torch.Tensor) -> int:
    n, c, h, w = t.shape
    return t.stride(3) if w > 1 else (t.stride(2) if h > 1 else (t.stride(0) if n > 1 else c))


def _nhwc(t: torch.Tensor) -> torch.Tensor:
    """A logical-NCHW fp32 tensor whose memory is NHWC with unit channel stride (a channels_last tensor or a channel
    slice of one) is passed through; anything else is copied into channels_last."""
    n, c, h, w = t.shape
    sn, sc, sh, sw = t.stride()
    ps = _pix_stride(t)
    ok = (t.dtype == torch.float32 and (sc == 1 or c == 1) and (w == 1 or h == 1 or sh == w * sw)
          and (n == 1 or sn == h * w * ps or (h == 1 and w == 1)) and ps % 4 == 0 and ps >= c and t.data_ptr() % 16 == 0)
    if ok:
        return t
    out = torch.empty((n, c, h, w), dtype=torch.float32, device=t.device, memory_format=CL)
    out.copy_(t)
    return out


def _desc(t: torch.Tensor) -> AddTensor:
    n, c, h, w = t.shape
    return AddTensor(t.data_ptr(), n, h, w, c, _pix_stride(t), ADD_F32)


def _new(n, c, h, w, dev) -> torch.Tensor:
    return torch.empty((n, c, h, w), dtype=torch.float32, device=dev, memory_format=CL)


def _ws(nbytes: int, dev) -> torch.Tensor:
    if nbytes < 0:
        check(int(nbytes), "workspace_bytes")
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)


# ---- conv ------------------------------------------------------------------------------------------------------------
class _Conv(torch.autograd.Function):
    """[ReLU ->] conv (dense, any k / stride / dilation), weights [Cout, Cin, kh, kw] like nn.Conv2d, optional bias.
    cout_pad: the output carries zero channels up to cout_pad (the 19-class classifier feeds 4-channel-vector kernels)."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, pad, dil, relu_in, cout_pad):
        x = _nhwc(x)
        n, cin, h, w = x.shape
        cout, cin_w, kh, kw = weight.shape
        cp = cout_pad or cout
        w_hwio = torch.zeros((kh, kw, cin, cp), dtype=torch.float32, device=x.device) if (cp != cout or cin != cin_w) else None
        if w_hwio is None:
            w_hwio = weight.detach().permute(2, 3, 1, 0).contiguous()
        else:
            w_hwio[:, :, :cin_w, :cout] = weight.detach().permute(2, 3, 1, 0)
        if pad >= 0:
            ho = (h + 2 * pad - dil * (kh - 1) - 1) // stride + 1
            wo = (w + 2 * pad - dil * (kw - 1) - 1) // stride + 1
        else:                                    # FactorizedReduce's shifted lattice (operations.py:97-99)
            ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
        y = _new(n, cp, ho, wo, x.device)
        b_ptr = None
        if bias is not None:
            bp = torch.zeros(cp, dtype=torch.float32, device=x.device)
            bp[:cout] = bias.detach()
            b_ptr = bp.data_ptr()
            ctx.bias_keep = bp
        check(lib.add_conv2d_fwd(ctypes.byref(_desc(x)), ctypes.byref(_desc(y)), w_hwio.data_ptr(), b_ptr, 0, kh, kw, stride, pad,
                                 dil, RELU_IN if relu_in else 0, _stream(x.device)), "train.conv_fwd")
        ctx.save_for_backward(x, w_hwio)
        ctx.cfg = (stride, pad, dil, relu_in, cout, cin_w, bias is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w_hwio = ctx.saved_tensors
        stride, pad, dil, relu_in, cout, cin_w, has_bias = ctx.cfg
        dy = _nhwc(dy)
        kh, kw, cin, cp = w_hwio.shape
        dev = x.device
        s = _stream(dev)
        n, _, h, w = x.shape
        dw = torch.empty_like(w_hwio)
        ws = _ws(lib.add_conv2d_wgrad_workspace_bytes(n, dy.shape[2], dy.shape[3], cin, cp, kh, kw), dev)
        check(lib.add_conv2d_wgrad(ctypes.byref(_desc(x)), ctypes.byref(_desc(dy)), dw.data_ptr(), kh, kw, stride, pad, dil,
                                   RELU_IN if relu_in else 0, ws.data_ptr(), ws.numel(), s), "train.conv_wgrad")
        dweight = dw[:, :, :cin_w, :cout].permute(3, 2, 0, 1)
        dbias = None
        if has_bias:
            dbias = _channel_sums(dy)[:cout]
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _new(n, cin, h, w, dev)
            if stride == 1 and pad >= 0:
                # dgrad of a stride-1 conv = the forward kernel with flipped, transposed weights and pad' = dil*(k-1) - pad
                wf = w_hwio.flip(0, 1).permute(0, 1, 3, 2).contiguous()
                check(lib.add_conv2d_fwd(ctypes.byref(_desc(dy)), ctypes.byref(_desc(dx)), wf.data_ptr(), None, 0, kh, kw, 1,
                                         dil * (kh - 1) - pad, dil, 0, s), "train.conv_dgrad_as_fwd")
            else:
                check(lib.add_conv2d_dgrad(ctypes.byref(_desc(dy)), w_hwio.data_ptr(), ctypes.byref(_desc(dx)), kh, kw, stride, pad,
                                           dil, 0, s), "train.conv_dgrad")
            if relu_in:
                check(lib.add_relu_mask_bwd(ctypes.byref(_desc(x)), ctypes.byref(_desc(dx)), s), "train.relu_mask")
        return dx, dweight, dbias, None, None, None, None, None


def conv2d(x, weight, bias=None, stride=1, pad=0, dil=1, relu_in=False, cout_pad=None):
    cout = weight.shape[0]
    if cout_pad is None and cout % 4:
        # the backward kernels read gradients as 4-channel vectors: carry zero channels up to a multiple of 4 and hand
        # out the channel slice (its gradient comes back zero-padded)
        return _Conv.apply(x, weight, bias, stride, pad, dil, relu_in, (cout + 3) // 4 * 4)[:, :cout]
    return _Conv.apply(x, weight, bias, stride, pad, dil, relu_in, cout_pad)


def _channel_sums(t: torch.Tensor) -> torch.Tensor:
    """sum over (n, h, w) per channel of an NHWC fp32 map: the deterministic GAP kernel, then N x C host-free glue."""
    n, c, h, w = t.shape
    pooled = torch.empty((n, c), dtype=torch.float32, device=t.device)
    ws = _ws(lib.add_global_avgpool_workspace_bytes(n, h, w, c), t.device)
    check(lib.add_global_avgpool_fwd(ctypes.byref(_desc(t)), pooled.data_ptr(), 0, ws.data_ptr(), ws.numel(), _stream(t.device)),
          "train.channel_sums")
    return pooled.sum(0) * float(h * w)


class _Depthwise(torch.autograd.Function):
    """[ReLU ->] depthwise k x k, stride 1, pad k//2; weight [C, 1, k, k] (operations.py:52,56)."""

    @staticmethod
    def forward(ctx, x, weight, relu_in):
        x = _nhwc(x)
        n, c, h, w = x.shape
        k = weight.shape[2]
        w_kkc = weight.detach()[:, 0].permute(1, 2, 0).contiguous()
        y = _new(n, c, h, w, x.device)
        check(lib.add_depthwise_fwd(ctypes.byref(_desc(x)), ctypes.byref(_desc(y)), w_kkc.data_ptr(), k, RELU_IN if relu_in else 0,
                                    _stream(x.device)), "train.dw_fwd")
        ctx.save_for_backward(x, w_kkc)
        ctx.relu_in = relu_in
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w_kkc = ctx.saved_tensors
        dy = _nhwc(dy)
        k = w_kkc.shape[0]
        n, c, h, w = x.shape
        dev, s = x.device, _stream(x.device)
        dw = torch.empty_like(w_kkc)
        ws = _ws(lib.add_depthwise_wgrad_workspace_bytes(n, h, w, c, k), dev)
        check(lib.add_depthwise_wgrad(ctypes.byref(_desc(x)), ctypes.byref(_desc(dy)), dw.data_ptr(), k, RELU_IN if ctx.relu_in else 0,
                                      ws.data_ptr(), ws.numel(), s), "train.dw_wgrad")
        dweight = dw.permute(2, 0, 1).unsqueeze(1)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _new(n, c, h, w, dev)
            wf = w_kkc.flip(0, 1).contiguous()
            check(lib.add_depthwise_fwd(ctypes.byref(_desc(dy)), ctypes.byref(_desc(dx)), wf.data_ptr(), k, 0, s), "train.dw_dgrad")
            if ctx.relu_in:
                check(lib.add_relu_mask_bwd(ctypes.byref(_desc(x)), ctypes.byref(_desc(dx)), s), "train.relu_mask")
        return dx, dweight, None


def depthwise(x, weight, relu_in=False):
    return _Depthwise.apply(x, weight, relu_in)


# ---- BatchNorm (training mode) -----------------------------------------------------------------------------------------
class _BatchNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, bn, relu, sync, group):
        x = _nhwc(x)
        n, c, h, w = x.shape
        dev, s = x.device, _stream(x.device)
        xd = _desc(x)
        stats = torch.empty(2 * c, dtype=torch.float32, device=dev)
        mean, inv_std = stats[:c], stats[c:]
        ws = _ws(lib.add_bn_stats_workspace_bytes(n, h, w, c), dev)
        packed = torch.empty(2 * c + 1, dtype=torch.float32, device=dev)
        check(lib.add_bn_stats_fwd(ctypes.byref(xd), packed.data_ptr(), ws.data_ptr(), ws.numel(), s), "train.bn_stats")
        packed[-1:].fill_(float(n * h * w))      # (a Python-scalar setitem is a host-to-device copy: not graph-capturable)
        if sync:
            sbn.exchange_sum(packed, group)
        if bn.num_batches_tracked is not None:
            bn.num_batches_tracked += 1
        momentum = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
        rm = bn.running_mean.data_ptr() if bn.track_running_stats else None
        rv = bn.running_var.data_ptr() if bn.track_running_stats else None
        check(lib.add_bn_finalize(packed.data_ptr(), packed.data_ptr() + 8 * c, 0.0, c, float(bn.eps), float(momentum),
                                  1 if sync else 0, rm, rv, mean.data_ptr(), inv_std.data_ptr(), s), "train.bn_finalize")
        y = _new(n, c, h, w, dev)
        check(lib.add_bn_apply_fwd(ctypes.byref(xd), ctypes.byref(_desc(y)), mean.data_ptr(), inv_std.data_ptr(),
                                   weight.data_ptr() if weight is not None else None, bias.data_ptr() if bias is not None else None,
                                   RELU_OUT if relu else 0, s), "train.bn_apply")
        var_term = None
        if sync:
            # inv_std = clamp(var, eps)^-1/2 (batchnorm.py:125): where the clamp is active the variance carries no gradient
            var_term = (inv_std < float(bn.eps) ** -0.5).to(torch.float32)
        ctx.save_for_backward(x, stats, weight, bias, packed, var_term)
        ctx.cfg = (relu, sync, group)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, stats, weight, bias, packed, var_term = ctx.saved_tensors
        relu, sync, group = ctx.cfg
        dy = _nhwc(dy)
        n, c, h, w = x.shape
        dev, s = x.device, _stream(x.device)
        mean, inv_std = stats[:c], stats[c:]
        sums = torch.empty(2 * c, dtype=torch.float64, device=dev)
        ws = _ws(lib.add_bn_bwd_workspace_bytes(n, h, w, c), dev)
        flags = RELU_OUT if relu else 0
        wp = weight.data_ptr() if weight is not None else None
        bp = bias.data_ptr() if bias is not None else None
        check(lib.add_bn_bwd_reduce(ctypes.byref(_desc(dy)), ctypes.byref(_desc(x)), mean.data_ptr(), inv_std.data_ptr(), wp, bp, flags,
                                    sums.data_ptr(), ws.data_ptr(), ws.numel(), s), "train.bn_bwd_reduce")
        dgamma = sums[c:].to(torch.float32) if weight is not None else None       # this rank's (DDP averages parameter grads)
        dbeta = sums[:c].to(torch.float32) if bias is not None else None
        if sync:
            sums = sums.clone()
            sbn.exchange_sum(sums, group)          # [sum dy, sum dy * xhat] over all ranks (batchnorm.py's autograd does the same)
        dx = _new(n, c, h, w, dev)
        # the element count of the (synchronised) batch is read on the device: packed[2C] holds it after the forward exchange
        check(lib.add_bn_bwd_apply(ctypes.byref(_desc(dy)), ctypes.byref(_desc(x)), mean.data_ptr(), inv_std.data_ptr(), wp, bp,
                                   sums.data_ptr(), 0.0, packed.data_ptr() + 8 * c, var_term.data_ptr() if var_term is not None else None,
                                   flags, ctypes.byref(_desc(dx)), s), "train.bn_bwd_apply")
        return dx, dgamma, dbeta, None, None, None, None


def batch_norm(bn: nn.BatchNorm2d, x, relu=False, sync: Optional[bool] = None, group=None):
    """Training-mode BatchNorm2d / SynchronizedBatchNorm2d with autograd.  sync None = synchronised iff the module is a
    SynchronizedBatchNorm2d with force_sync or a process group of more than one rank is initialised (the reference's
    synchronised formulas: clamp(var, eps), batchnorm.py:113-125); else F.batch_norm's (var + eps)."""
    import torch.distributed as dist
    if sync is None:
        sync = bool(getattr(bn, "force_sync", False)) or (isinstance(bn, sbn.SynchronizedBatchNorm2d) and dist.is_available()
                                                          and dist.is_initialized() and dist.get_world_size(group) > 1)
    group = group if group is not None else getattr(bn, "process_group", None)
    return _BatchNorm.apply(x, bn.weight if bn.affine else None, bn.bias if bn.affine else None, bn, relu, sync, group)


# ---- bilinear ---------------------------------------------------------------------------------------------------------
_BIL_TABLES: Dict[tuple, tuple] = {}


def _bil_tables(in_size: int, out_size: int, dev) -> tuple:
    key = (in_size, out_size, str(dev))
    t = _BIL_TABLES.get(key)
    if t is None:
        i0 = torch.empty(out_size, dtype=torch.int32, device=dev); i1 = torch.empty_like(i0)
        l0 = torch.empty(out_size, dtype=torch.float32, device=dev); l1 = torch.empty_like(l0)
        lo = torch.empty(in_size, dtype=torch.int32, device=dev); hi = torch.empty_like(lo)
        check(lib.add_bilinear_bwd_tables(in_size, out_size, i0.data_ptr(), i1.data_ptr(), l0.data_ptr(), l1.data_ptr(),
                                          lo.data_ptr(), hi.data_ptr(), _stream(dev)), "train.bilinear_tables")
        t = _BIL_TABLES[key] = (i0, i1, l0, l1, lo, hi)
    return t


class _Bilinear(torch.autograd.Function):
    """F.interpolate(x, size, mode='bilinear') (align_corners=False; ADD.py:76,84,89,317; decoder.py:24,28)."""

    @staticmethod
    def forward(ctx, x, ho, wo):
        x = _nhwc(x)
        n, c, h, w = x.shape
        y = _new(n, c, ho, wo, x.device)
        check(lib.add_bilinear_fwd(ctypes.byref(_desc(x)), ctypes.byref(_desc(y)), 0, _stream(x.device)), "train.bilinear_fwd")
        ctx.shape = (n, c, h, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _nhwc(dy)
        n, c, h, w = ctx.shape
        dev = dy.device
        dx = _new(n, c, h, w, dev)
        ty, tx = _bil_tables(h, dy.shape[2], dev), _bil_tables(w, dy.shape[3], dev)
        check(lib.add_bilinear_bwd(ctypes.byref(_desc(dy)), ctypes.byref(_desc(dx)), *[t.data_ptr() for t in ty],
                                   *[t.data_ptr() for t in tx], 0, _stream(dev)), "train.bilinear_bwd")
        return dx, None, None


def bilinear(x, size):
    if x.shape[2] == size[0] and x.shape[3] == size[1]:
        return x            # F.interpolate to the same size is the identity for align_corners=False
    return _Bilinear.apply(x, int(size[0]), int(size[1]))


# ---- glue with our copy kernel: concat / add / global pooling / broadcast -----------------------------------------------
def _copy_into(src, dst, accumulate=False, scale=1.0):
    check(lib.add_scale_fwd(ctypes.byref(_desc(src)), ctypes.byref(_desc(dst)), float(scale), 1, ACCUMULATE if accumulate else 0,
                            _stream(src.device)), "train.copy")


class _Cat(torch.autograd.Function):
    """torch.cat(dim=1) as slice writes into one NHWC buffer; the backward hands out channel-slice views (no copy)."""

    @staticmethod
    def forward(ctx, *xs):
        xs = [_nhwc(x) for x in xs]
        n, _, h, w = xs[0].shape
        cs = [x.shape[1] for x in xs]
        y = _new(n, sum(cs), h, w, xs[0].device)
        off = 0
        for x, c in zip(xs, cs):
            _copy_into(x, y[:, off:off + c])
            off += c
        ctx.cs = cs
        return y

    @staticmethod
    def backward(ctx, dy):
        out, off = [], 0
        for c in ctx.cs:
            out.append(dy[:, off:off + c])
            off += c
        return tuple(out)


def cat(xs: Sequence[torch.Tensor]):
    return _Cat.apply(*xs)


class _Add(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = _nhwc(a), _nhwc(b)
        y = _new(*a.shape, a.device)
        _copy_into(a, y)
        _copy_into(b, y, accumulate=True)
        return y

    @staticmethod
    def backward(ctx, dy):
        return dy, dy


def add(a, b):
    return _Add.apply(a, b)


class _GlobalAvgPool(torch.autograd.Function):
    """nn.AdaptiveAvgPool2d(1): [N,C,H,W] -> [N,C,1,1]; backward broadcasts g / (H*W)."""

    @staticmethod
    def forward(ctx, x, relu_in):
        x = _nhwc(x)
        n, c, h, w = x.shape
        out = torch.empty((n, c), dtype=torch.float32, device=x.device)
        ws = _ws(lib.add_global_avgpool_workspace_bytes(n, h, w, c), x.device)
        check(lib.add_global_avgpool_fwd(ctypes.byref(_desc(x)), out.data_ptr(), RELU_IN if relu_in else 0, ws.data_ptr(), ws.numel(),
                                         _stream(x.device)), "train.gap")
        ctx.save_for_backward(x)
        ctx.relu_in = relu_in
        return out.view(n, c, 1, 1).contiguous(memory_format=CL)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        n, c, h, w = x.shape
        g = dy.reshape(n, c, 1, 1).contiguous(memory_format=CL)
        dx = _new(n, c, h, w, x.device)
        # broadcast the 1x1 gradient (bilinear from a 1-pixel source is an exact copy), then scale by 1 / (H*W) in place
        check(lib.add_bilinear_fwd(ctypes.byref(_desc(g)), ctypes.byref(_desc(dx)), 0, _stream(x.device)), "train.gap_bwd_broadcast")
        _copy_into(dx, dx, scale=1.0 / float(h * w))
        if ctx.relu_in:
            check(lib.add_relu_mask_bwd(ctypes.byref(_desc(x)), ctypes.byref(_desc(dx)), _stream(x.device)), "train.relu_mask")
        return dx, None


class _Broadcast(torch.autograd.Function):
    """nn.Upsample(size, bilinear, align_corners=True) of a 1x1 map (aspp_train.py:54-55) = a broadcast; the backward sums
    over the pixels."""

    @staticmethod
    def forward(ctx, x, ho, wo):
        x = _nhwc(x)
        n, c = x.shape[:2]
        y = _new(n, c, ho, wo, x.device)
        check(lib.add_bilinear_fwd(ctypes.byref(_desc(x)), ctypes.byref(_desc(y)), 0, _stream(x.device)), "train.broadcast")
        return y

    @staticmethod
    def backward(ctx, dy):
        dy = _nhwc(dy)
        n, c, h, w = dy.shape
        pooled = torch.empty((n, c), dtype=torch.float32, device=dy.device)
        ws = _ws(lib.add_global_avgpool_workspace_bytes(n, h, w, c), dy.device)
        check(lib.add_global_avgpool_fwd(ctypes.byref(_desc(dy)), pooled.data_ptr(), 0, ws.data_ptr(), ws.numel(), _stream(dy.device)),
              "train.broadcast_bwd")
        return (pooled * float(h * w)).view(n, c, 1, 1), None, None


# ---- loss ---------------------------------------------------------------------------------------------------------------
class _CrossEntropy(torch.autograd.Function):
    """nn.CrossEntropyLoss(weight, ignore_index=255) on [N, C(+pad), H, W] logits (utils/loss.py:16-25)."""

    @staticmethod
    def forward(ctx, logits, target, num_class, ignore_index, class_weight):
        n, cp, h, w = logits.shape
        x = logits.permute(0, 2, 3, 1)                       # NHWC view of the channels_last logits
        assert x.is_contiguous() or logits.is_contiguous(), "logits must be channels_last or contiguous NCHW"
        nhwc = x.is_contiguous()
        out = torch.empty(2, dtype=torch.float32, device=logits.device)
        dlog = torch.empty_like(logits)
        ws = _ws(lib.add_ce_loss_workspace_bytes(n, h, w), logits.device)
        tgt = target.to(torch.int64).contiguous()
        sc, sp, sn = (1, cp, h * w * cp) if nhwc else (h * w, 1, cp * h * w)
        check(lib.add_ce_loss_fwd_bwd(logits.data_ptr(), tgt.data_ptr(), n, num_class, h, w, sn, sc, sp, cp, int(ignore_index),
                                      class_weight.data_ptr() if class_weight is not None else None, 1.0, out.data_ptr(),
                                      dlog.data_ptr(), ws.data_ptr(), ws.numel(), _stream(logits.device)), "train.ce")
        ctx.save_for_backward(dlog)
        return out[0].clone()

    @staticmethod
    def backward(ctx, g):
        (dlog,) = ctx.saved_tensors
        return dlog * g, None, None, None, None


def cross_entropy(logits, target, num_class=19, ignore_index=255, class_weight=None):
    return _CrossEntropy.apply(logits, target, num_class, ignore_index, class_weight)


# ---- the reference modules in train mode ----------------------------------------------------------------------------------
def relu_conv_bn(m, x):
    """operations.py:18-29."""
    conv = m.op[1]
    y = conv2d(x, conv.weight, None, conv.stride[0], conv.padding[0], conv.dilation[0], relu_in=True)
    return batch_norm(m.op[2], y)


def dil_conv(m, x):
    """operations.py:32-43 (dense dilated conv)."""
    conv = m.op[1]
    y = conv2d(x, conv.weight, None, conv.stride[0], conv.padding[0], conv.dilation[0], relu_in=True)
    return batch_norm(m.op[2], y)


def sep_conv(m, x):
    """operations.py:46-62."""
    y = depthwise(x, m.op[1].weight, relu_in=True)
    y = conv2d(y, m.op[2].weight)
    y = batch_norm(m.op[3], y, relu=True)            # BN -> ReLU (op.4) fused
    y = depthwise(y, m.op[5].weight)
    y = conv2d(y, m.op[6].weight)
    return batch_norm(m.op[7], y)


def factorized_reduce(m, x):
    """operations.py:86-101 / :104-119: relu; cat(conv_1(x), conv_2(pad(x)[:, :, 1:, 1:])); bn."""
    st = m.STEP
    a = conv2d(x, m.conv_1.weight, None, st, 0, 1, relu_in=True)
    b = conv2d(x, m.conv_2.weight, None, st, -(st // 2), 1, relu_in=True)
    return batch_norm(m.bn, cat([a, b]))


def _op_forward(op, x):
    from .operations import SepConv, DilConv, Identity, Zero, _Pool3x3
    if isinstance(op, SepConv):
        return sep_conv(op, x)
    if isinstance(op, DilConv):
        return dil_conv(op, x)
    if isinstance(op, Identity):
        return x
    raise NotImplementedError(f"training forward of {type(op).__name__} (pools / none are supernet primitives: cell_level_search)")


def _prep(m, x):
    from .operations import _FactorizedReduceBase
    return factorized_reduce(m, x) if isinstance(m, _FactorizedReduceBase) else relu_conv_bn(m, x)


def cell_forward(cell, prev_prev, prev):
    """ADD.py:69-116."""
    s1 = prev
    if cell.downup_sample == 1:
        s1 = bilinear(s1, (cell.scale_dimension(s1.shape[2], cell.scale), cell.scale_dimension(s1.shape[3], cell.scale)))
    s1 = _prep(cell.preprocess, s1)
    size = (s1.shape[2], s1.shape[3])
    if not cell.dense_in:
        s0 = prev_prev if prev_prev.shape[2] == size[0] else bilinear(prev_prev, size)
        s0 = relu_conv_bn(cell.pre_preprocess, s0)
    else:
        parts = []
        for i, t in enumerate(prev_prev):
            t = t if t.shape[2] == size[0] else bilinear(t, size)
            parts.append(relu_conv_bn(cell.pre_preprocess[i], t))
        s0 = relu_conv_bn(cell.pre_preprocess_1x1, cat(parts))
    states = [s0, s1]
    for edges in cell._steps:
        new = [_op_forward(cell._ops[k], states[j]) for (j, k) in edges]
        s = new[0]
        for t in new[1:]:
            s = add(s, t)
        states.append(s)
    concat = cat(states[-cell.B:])
    if cell.dense_out:
        return prev, concat, relu_conv_bn(cell.dense_process, concat)
    return concat


def aspp_forward(m, x):
    """aspp_train.py:34-61."""
    h, w = x.shape[2], x.shape[3]
    outs = []
    for k in range(1, 5):
        conv = getattr(m, f"aspp{k}")
        y = conv2d(x, conv.weight, None, 1, conv.padding[0], conv.dilation[0], relu_in=True)
        outs.append(batch_norm(getattr(m, f"aspp{k}_bn"), y, relu=True))
    p = _GlobalAvgPool.apply(x, True)
    p = batch_norm(m.aspp5_bn, conv2d(p, m.aspp5.weight), relu=True)
    outs.append(_Broadcast.apply(p, h, w))
    y = conv2d(cat(outs), m.conv1.weight)
    return batch_norm(m.bn1, y)


def decoder_forward(m, x, low_level, size):
    """decoder.py:23-29; the classifier output is carried with one zero channel (19 -> 20) through the final bilinear."""
    if x.shape[2] != low_level.shape[2]:
        x = bilinear(x, (low_level.shape[2], low_level.shape[3]))
    y = cat([x, low_level])
    y = batch_norm(m._conv[2], conv2d(y, m._conv[1].weight, None, 1, 1, 1, relu_in=True), relu=True)
    y = batch_norm(m._conv[5], conv2d(y, m._conv[4].weight, None, 1, 1, 1), relu=True)
    nc = m._conv[7].weight.shape[0]
    y = conv2d(y, m._conv[7].weight, m._conv[7].bias, cout_pad=(nc + 3) // 4 * 4)
    return bilinear(y, size)


def add_forward(net, x: torch.Tensor) -> List[torch.Tensor]:
    """ADD.forward (ADD.py:277-325) in training mode.  x: [N,3,H,W] fp32 CUDA.  Returns the C logits tensors
    [N, 20, H, W] (channels_last; channel 19 is padding and is ignored by `cross_entropy`)."""
    n, _, H, W = x.shape
    size = (H, W)
    sc = 2.0 ** (-1 * (net.network_arch[-1] + 2))
    aspp_size = (int((float(H) - 1.0) * sc + 1.0), int((float(W) - 1.0) * sc + 1.0))
    img = torch.zeros((n, 4, H, W), dtype=torch.float32, device=x.device).contiguous(memory_format=CL)
    img[:, :3] = x
    t = batch_norm(net.stem0[1], conv2d(img, net.stem0[0].weight, None, 2, 1, 1), relu=True)
    # stem2's in-place ReLU acts on stem1's output, which cell 0 also reads (Q7): every reader sees relu(stem1 output)
    stem0 = batch_norm(net.stem1[1], conv2d(t, net.stem1[0].weight, None, 1, 1, 1), relu=True)
    stem1 = batch_norm(net.stem2[2], conv2d(stem0, net.stem2[1].weight, None, 2, 1, 1))
    two = [stem0, stem1]
    dense, outs, it = [], [], 0
    cur = None
    low_level = None
    for i in range(net.num_net):
        cell = net.cells[i]
        if i < 3:
            two[0], two[1], fm = cell_forward(cell, two[0], two[1])
            dense.append(fm)
            if i == 2:
                cur = two[1]
        elif i < net.num_net - 2:
            _, cur, fm = cell_forward(cell, list(dense[:-1]), cur)
            dense.append(fm)
        elif i == net.num_net - 1:
            cur = cell_forward(cell, list(dense), cur)
        else:
            cur = cell_forward(cell, list(dense[:-1]), cur)
        if i == net.low_level_layer:
            conv = net.low_level_conv[1]
            low_level = batch_norm(net.low_level_conv[2], conv2d(two[1], conv.weight, relu_in=True))
        if i in net.C_index or i == net.num_net - 1:
            y = cur if i > 2 else two[1]
            if y.shape[2] < aspp_size[0] or y.shape[3] < aspp_size[1]:
                y = bilinear(y, aspp_size)
            if net.network_arch[i] != net.network_arch[-1]:
                y = _prep(net.conv_aspp[it], y)
                it += 1
            y = aspp_forward(net.aspp, y)
            outs.append(decoder_forward(net.decoder, y, low_level, size))
    return outs


def add_loss(net, x, target, class_weight=None):
    """train.py:227-233: mean over the C exits of CrossEntropy(ignore 255)."""
    outs = add_forward(net, x)
    losses = [cross_entropy(o, target, net._num_classes, 255, class_weight) for o in outs]
    total = losses[0]
    for l in losses[1:]:
        total = total + l
    return total / float(len(losses)), outs


# ---- optimizer ---------------------------------------------------------------------------------------------------------
class SGD:
    """torch.optim.SGD(params, lr, momentum, weight_decay, nesterov) (train.py:126-127) as ONE kernel launch over a
    device table of (param, grad, momentum buffer) pointers; gradients live in one flat buffer (so a data-parallel step
    all-reduces them with one collective)."""

    def __init__(self, params, lr, momentum=0.9, weight_decay=4e-5, nesterov=True):
        self.params = [p for p in params if p.requires_grad]
        self.lr, self.momentum, self.weight_decay, self.nesterov = float(lr), float(momentum), float(weight_decay), bool(nesterov)
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_buf = torch.zeros(total, dtype=torch.float32, device=dev)
        off, rows = 0, []
        for p in self.params:
            n = p.numel()
            assert p.is_contiguous() and p.dtype == torch.float32
            p.grad = self.flat_grad[off:off + n].view_as(p)
            rows.append((p.data_ptr(), self.flat_grad.data_ptr() + 4 * off, self.flat_buf.data_ptr() + 4 * off, n))
            off += n
        import numpy as np
        self.table = torch.from_numpy(np.array(rows, dtype=np.int64)).to(dev)
        self.max_numel = max(r[3] for r in rows)
        self.steps = 0
        self.lr_dev = torch.full((1,), self.lr, dtype=torch.float32, device=dev)   # read by the kernel (graph-replayable)

    def zero_grad(self):
        self.flat_grad.zero_()

    def all_reduce_grads(self, group=None):
        """DDP's gradient averaging (train.py:173) as one all-reduce of the flat gradient buffer."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=group)
            self.flat_grad.mul_(1.0 / dist.get_world_size(group))

    def set_lr(self, lr: float) -> None:
        self.lr = float(lr)
        self.lr_dev.fill_(self.lr)

    def step(self, lr: Optional[float] = None):
        if lr is not None and float(lr) != self.lr:
            self.set_lr(lr)
        check(lib.add_sgd_nesterov(self.table.data_ptr(), len(self.params), self.max_numel, self.lr, self.lr_dev.data_ptr(), self.momentum,
                                   self.weight_decay, 1 if self.nesterov else 0, 1 if self.steps == 0 else 0,
                                   _stream(self.flat_grad.device)), "train.sgd")
        self.steps += 1
        from . import runtime as rt
        rt.bump_generation()                     # parameters moved: folded eval-mode weights / recorded plans are stale


def poly_lr(base_lr: float, it: int, total_iters: int, power: float = 0.9, min_lr: Optional[float] = None) -> float:
    """utils/lr_scheduler.py:50-51 ('poly')."""
    lr = base_lr * pow((1 - 1.0 * it / total_iters), power)
    if min_lr is not None and lr < min_lr:
        lr = min_lr
    return lr


def train_step(net, optimizer: SGD, x, target, lr=None, class_weight=None, group=None):
    """One iteration of train.py:216-247.  Returns the loss (a device scalar)."""
    if lr is not None:
        optimizer.set_lr(lr)
    optimizer.zero_grad()
    loss, _ = add_loss(net, x, target, class_weight)
    loss.backward()
    optimizer.all_reduce_grads(group)
    optimizer.step()
    return loss.detach()


class GraphedTrainStep:
    """The whole iteration — forward, loss, backward, SyncBN exchanges, gradient all-reduce, SGD — as ONE CUDA graph.
    Eagerly the step is ~5000 kernel launches issued from Python (host-bound: ~250 ms whatever the GPU does); captured,
    the launches replay back to back.  The first `warmup` calls run eagerly (they are real training steps; they also build
    the bilinear tables and the peer-exchange buffers), the next call captures the step and from then on every call copies
    its batch into the static input buffers, sets the learning rate on the device and replays.  Everything that changes
    between iterations is read on the device: the batch, lr (`add_sgd_nesterov`'s lr_dev), the peer exchange's call counter."""

    def __init__(self, net, optimizer: SGD, class_weight=None, group=None, warmup: int = 2):
        self.net, self.opt, self.cw, self.group, self.warmup = net, optimizer, class_weight, group, max(1, int(warmup))
        self.calls = 0
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.x = self.gt = self.loss = None

    def __call__(self, x: torch.Tensor, target: torch.Tensor, lr: Optional[float] = None) -> torch.Tensor:
        self.calls += 1
        if self.calls <= self.warmup:
            return train_step(self.net, self.opt, x, target, lr, self.cw, self.group)
        if lr is not None:
            self.opt.set_lr(lr)
        if self.graph is None:
            self.x, self.gt = x.clone(), target.clone()
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.loss = train_step(self.net, self.opt, self.x, self.gt, None, self.cw, self.group)
        else:
            self.x.copy_(x, non_blocking=True)
            self.gt.copy_(target, non_blocking=True)
        self.graph.replay()
        self.opt.steps += 1
        return self.loss
