"""TEST INFRASTRUCTURE ONLY — a CPU stand-in for the launch layer, for host-logic tests without a GPU.

The product has no CPU path and must not get one: nothing under `auto-dynamic-deeplab_b200/` knows this file exists, and
nothing here is importable from the product.  `install(monkeypatch)` replaces, for the duration of ONE test, the op methods
of `add_b200.runtime.Builder` (the single place where the Python side turns "conv this view into that channel slice,
accumulate, ReLU on store" into a C-ABI launch) by plain-PyTorch CPU closures with the semantics `include/add_b200.h`
documents for each entry point.  Everything above that line runs unchanged: plan recording, channel-slice concat,
accumulate-into-slice node sums, ReLU-on-load/store bookkeeping (`relud`), resized-feature sharing, early-exit segments,
image gathers driven by index vectors the host fills in later, per-plan result scatter.  So a wiring mistake in that host
logic — a wrong slice offset, a missing ReLU flag, a segment continuing from the wrong state buffers — shows up as a
numerical mismatch against the reference's golden outputs right here, on CPU.

What it does NOT test: the kernels (their parity is the `-m gpu` suite) and the C-side argument marshalling
(tests/test_dry_run_plans.py hands every recorded launch to the real entry points for validation)."""
from __future__ import annotations

import contextlib

import torch
import torch.nn.functional as F

from add_b200 import runtime as rt
from add_b200.runtime import Builder, Plan, View, RELU_IN, RELU_OUT, ACCUMULATE


def _sl(v: View) -> torch.Tensor:
    return v.buf[..., v.c_off:v.c_off + v.c]


def _x(v: View, relu) -> torch.Tensor:
    t = _sl(v).float().permute(0, 3, 1, 2)
    return torch.relu(t) if relu else t


def _store(v: View, out_nchw: torch.Tensor, flags: int) -> None:
    dst = _sl(v)
    o = out_nchw.permute(0, 2, 3, 1)
    assert tuple(o.shape) == tuple(dst.shape), (tuple(o.shape), tuple(dst.shape))
    if flags & ACCUMULATE:
        o = o + dst.float()
    if flags & RELU_OUT:
        o = torch.relu(o)
    dst.copy_(o.to(dst.dtype))


def _do(b: Builder, run, tag: str, kernel: str) -> None:
    if b.record:
        def launch(_stream=None):
            run()
            return 0
        b.launches.append((launch, (), tag, dict(kernel=kernel, flops=0, bytes=0, reads=[], writes=[])))
    else:
        run()


def _window(xin: torch.Tensor, out_h: int, out_w: int, k_h: int, k_w: int, stride: int, pad: int, dil: int) -> torch.Tensor:
    """The input region the conv reads for exactly out_h x out_w outputs, taps outside the image as zeros (pad may be negative)."""
    n, c, H, W = xin.shape
    need_h, need_w = (out_h - 1) * stride + dil * (k_h - 1) + 1, (out_w - 1) * stride + dil * (k_w - 1) + 1
    xp = xin.new_zeros((n, c, need_h, need_w))
    y0, x0 = -pad, -pad                                  # input coordinate of xp[.., 0, 0]
    ys, ye = max(0, y0), min(H, y0 + need_h)
    xs, xe = max(0, x0), min(W, x0 + need_w)
    if ye > ys and xe > xs:
        xp[:, :, ys - y0:ye - y0, xs - x0:xe - x0] = xin[:, :, ys:ye, xs:xe]
    return xp


def conv(self, x, y, cw, stride=1, pad=0, dil=1, flags=0, tag="conv", image_bias=None):
    assert x.c == cw.cin and y.c == cw.cout and x.n == y.n, (x.c, cw.cin, y.c, cw.cout, tag)
    if x.relud:
        assert flags & RELU_IN, f"{tag}: a post-ReLU buffer read by a conv that does not start with ReLU"
        flags &= ~RELU_IN
    self.keep.append(cw)

    def run():
        xin = _window(_x(x, flags & RELU_IN), y.h, y.w, cw.kh, cw.kw, stride, pad, dil)
        w = cw.w_h.permute(3, 2, 0, 1).contiguous()
        if x.dtype == torch.bfloat16:            # the tcgen05 path multiplies bf16 weights (fp32 accumulate)
            w = w.to(torch.bfloat16).float()
        out = F.conv2d(xin, w, None, stride, 0, dil)
        if image_bias is not None:
            out = out + image_bias.float().view(x.n, cw.cout, 1, 1)
        elif cw.bias is not None:
            out = out + cw.bias.float().view(1, -1, 1, 1)
        _store(y, out, flags)
    _do(self, run, tag, "conv2d")


def sepconv_half(self, x, y, w_dw, pw, k, flags, tag="sephalf"):
    if x.relud and (flags & RELU_IN):
        flags &= ~RELU_IN
    assert (x.n, x.h, x.w) == (y.n, y.h, y.w) and pw.cin == x.c and pw.cout == y.c

    def run():
        xin = _x(x, flags & RELU_IN)
        d = F.conv2d(xin, w_dw.float().permute(2, 0, 1).unsqueeze(1).contiguous(), None, 1, k // 2, 1, x.c)
        w = pw.w_h.permute(3, 2, 0, 1).contiguous()
        if x.dtype == torch.bfloat16:            # depthwise result rounded to bf16 (the UMMA A tile), bf16 pointwise weights
            d, w = d.to(torch.bfloat16).float(), w.to(torch.bfloat16).float()
        out = F.conv2d(d, w)
        if pw.bias is not None:
            out = out + pw.bias.float().view(1, -1, 1, 1)
        _store(y, out, flags)
    _do(self, run, tag, "sepconv_half")


def pool3x3(self, x, y, mode, stride, flags=0, tag="pool3x3"):
    assert not x.relud, f"{tag}: raw read of a buffer that was stored post-ReLU"

    def run():
        xin = _x(x, False)
        out = (F.max_pool2d(xin, 3, stride, 1) if mode == 1 else F.avg_pool2d(xin, 3, stride, 1, count_include_pad=False))
        _store(y, out, flags)
    _do(self, run, tag, "pool3x3")


def scale(self, x, y, scale, stride=1, flags=0, tag="scale"):
    assert not x.relud, f"{tag}: raw read of a buffer that was stored post-ReLU"
    _do(self, lambda: _store(y, float(scale) * _x(x, False)[:, :, ::stride, ::stride], flags), tag, "scale")


def bilinear(self, x, y, flags=0, tag="bilinear"):
    assert not x.relud or (flags & RELU_IN), f"{tag}: raw read of a buffer that was stored post-ReLU"

    def run():
        xin = _x(x, flags & RELU_IN)
        out = xin if (x.h, x.w) == (y.h, y.w) else F.interpolate(xin, (y.h, y.w), mode="bilinear", align_corners=False)
        _store(y, out, flags & RELU_OUT)
    _do(self, run, tag, "bilinear")


def gather_images(self, src, dst, idx, tag="gather_images"):
    assert src.shape[1:] == dst.shape[1:] and src.dtype == dst.dtype and idx.dtype == torch.int32
    _do(self, lambda: dst.copy_(src[idx[:dst.shape[0]].long()]), tag, "gather_images")


def gather_view(self, src, dst, idx, tag="gather_view"):
    assert (src.h, src.w, src.c, src.dtype) == (dst.h, dst.w, dst.c, dst.dtype) and idx.dtype == torch.int32
    _do(self, lambda: _sl(dst).copy_(_sl(src)[idx[:dst.n].long()]), tag, "gather_images")


def gap(self, x, out, flags=0, tag="gap"):
    _do(self, lambda: out.copy_(_x(x, (flags & RELU_IN) and not x.relud).mean(dim=(2, 3))), tag, "global_avgpool")


def aspp_pool_bias(self, pooled, cw5, w_out_pool, b_out, out, tag="aspp_pool_bias"):
    def run():
        h = pooled.float() @ cw5.w_h[0, 0]
        if cw5.bias is not None:
            h = h + cw5.bias.float()
        o = torch.relu(h) @ w_out_pool.float()
        out.copy_(o + b_out.float() if b_out is not None else o)
    _do(self, run, tag, "aspp_pool_bias")


def nchw_to_nhwc(self, src, c_src, y, tag="nchw2nhwc"):
    def run():
        dst = _sl(y)
        dst.zero_()
        dst[..., :c_src] = src.permute(0, 2, 3, 1).to(dst.dtype)
    _do(self, run, tag, "nchw_to_nhwc")


def nhwc_to_nchw(self, x, dst, tag="nhwc2nchw"):
    _do(self, lambda: dst.copy_(_x(x, False)), tag, "nhwc_to_nchw")


def _up(x: View, H: int, W: int) -> torch.Tensor:
    return F.interpolate(_x(x, False), (H, W), mode="bilinear", align_corners=False)


def upsample_logits(self, x, dst, H, W, tag="upsample_logits"):
    _do(self, lambda: dst.copy_(_up(x, H, W)), tag, "upsample_logits_nchw")


def upsample_argmax(self, x, H, W, gt, pred, cm, ent, tag="upsample_argmax", cm_rows=None):
    nc = x.c

    def run():
        lg = _up(x, H, W)
        p = lg.argmax(1)
        if pred is not None:
            pred.copy_(p)
        if cm is not None:
            for j in range(x.n):
                g = gt[j].reshape(-1).long()
                q = p[j].reshape(-1)
                keep = (g >= 0) & (g < nc)
                row = int(cm_rows[j]) if cm_rows is not None else j
                cm[row] = torch.bincount(nc * g[keep] + q[keep], minlength=nc * nc).reshape(nc, nc)
        if ent is not None:
            lp = torch.log_softmax(lg, 1)
            e = -(lp.exp() * lp).sum(1) / torch.log(torch.tensor(float(nc)))
            ent.copy_(e.sum(dim=(1, 2)) / float(H * W))
    _do(self, run, tag, "upsample_argmax")


def edm_mlp(self, pooled, n, ws, out, tag="edm_mlp"):
    def run():
        w0, b0, w1, b1, w2, b2 = [t.float() for t in ws]
        h = torch.relu(pooled[:n].float() @ w0.t() + b0)
        h = torch.relu(h @ w1.t() + b1)
        out.copy_((h @ w2.t() + b2).reshape(-1))
    _do(self, run, tag, "EDM.mlp")


def stem_nchw(self, src, y, w_packed, bias, flags, tag="stem0"):
    """The fused bf16 stem (NCHW fp32 image -> bf16 -> 3x3 s2 conv 3->64 + bias -> ReLU -> NHWC bf16).  `w_packed` is the
    ConvWeights itself here: install() makes `rt.pack_stem_tc` the identity, the UMMA image is not decoded on CPU."""
    cw = w_packed

    def run():
        xin = _window(src.to(torch.bfloat16).float(), y.h, y.w, 3, 3, 2, 1, 1)
        w = cw.w_h[:, :, :3, :].to(torch.bfloat16).float().permute(3, 2, 0, 1).contiguous()
        out = F.conv2d(xin, w, None, 2) + bias.float().view(1, -1, 1, 1)
        _store(y, out, flags)
    _do(self, run, tag, "stem_conv_tc")


class _Stream:
    cuda_stream = 0

    def wait_stream(self, *_):
        pass

    def wait_event(self, *_):
        pass

    def synchronize(self):
        pass


class _Event:
    def __init__(self, *a, **k):
        pass

    def record(self, *_):
        pass

    def synchronize(self):
        pass

    def wait(self, *_):
        pass


def install(monkeypatch) -> None:
    """Patch the launch layer for one test (pytest's monkeypatch undoes everything afterwards)."""
    import add_b200.dynamic as dyn
    from add_b200.ADD import _NetPlan
    for name, fn in dict(conv=conv, sepconv_half=sepconv_half, pool3x3=pool3x3, scale=scale, bilinear=bilinear,
                         gather_images=gather_images, gather_view=gather_view, gap=gap, aspp_pool_bias=aspp_pool_bias,
                         nchw_to_nhwc=nchw_to_nhwc, nhwc_to_nchw=nhwc_to_nchw, upsample_logits=upsample_logits,
                         upsample_argmax=upsample_argmax, edm_mlp=edm_mlp, stem_nchw=stem_nchw).items():
        monkeypatch.setattr(Builder, name, fn)

    def run_eager(self):
        for fn, args, tag, _ in self.launches:
            assert fn(*args, None) == 0, tag
    monkeypatch.setattr(Plan, "run_eager", run_eager)
    monkeypatch.setattr(rt, "require_cuda", lambda *a, **k: None)
    monkeypatch.setattr(rt, "require_cuda_device", lambda *a, **k: None)
    monkeypatch.setattr(rt, "pack_stem_tc", lambda cw: cw)
    monkeypatch.setattr(torch.cuda, "Stream", lambda *a, **k: _Stream())
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())

    def fresh_logits(self):
        n, nc, H, W = self.out_shape
        return [_up(lg, H, W).clone() for lg in self.lowres]

    def fresh_feature(self):
        return _x(self.feature, False).clone()
    monkeypatch.setattr(_NetPlan, "fresh_logits", fresh_logits)
    monkeypatch.setattr(_NetPlan, "fresh_feature", fresh_feature)
    monkeypatch.setattr(dyn, "_OVERLAP_HEADS", False)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _Stream())
    monkeypatch.setattr(torch.cuda, "Event", _Event)
