"""Host-side runtime: NHWC views, folded weights, and the launch recorder / replayer.

The drop-in modules never call PyTorch compute ops on activations.  They *emit* libadd_b200
launches through a `Builder`: eagerly (per-op modules called on their own) or recorded into a
`Plan` that `ADD.forward` replays — optionally as one CUDA graph — with every concat / node-sum
expressed as channel-slice writes.  PyTorch supplies device memory, streams and graphs only.
"""
from __future__ import annotations

import ctypes
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import AddTensor, lib, check, ADD_F32, ADD_BF16, RELU_IN, RELU_OUT, ACCUMULATE

IN_RELUD = 8   # host-only emit flag: the input view already holds relu(x) (its producer stored it with RELU_OUT), so
               # the op skips its ReLU-on-load.  Exact: relu(round(x)) == round(relu(x)).  Never passed to the C ABI.

_GENERATION = 0  # bumped whenever any add_b200 module's parameters may have changed


def bump_generation() -> None:
    global _GENERATION
    _GENERATION += 1


def generation() -> int:
    return _GENERATION


_DTYPES = {torch.float32: ADD_F32, torch.bfloat16: ADD_BF16}
_PRECISION = {"fp32": torch.float32, "bf16": torch.bfloat16}
_default_precision = "fp32"


def set_default_precision(p: str) -> None:
    """'fp32' (exact CUDA-core path) or 'bf16' (bf16 activations, tensor-core contractions)."""
    global _default_precision
    if p not in _PRECISION:
        raise ValueError(f"precision must be one of {list(_PRECISION)}")
    _default_precision = p
    bump_generation()


def default_precision() -> str:
    return _default_precision


def act_dtype(precision: Optional[str] = None) -> torch.dtype:
    return _PRECISION[precision or _default_precision]


def require_cuda(t: torch.Tensor, what: str = "input") -> None:
    if not t.is_cuda:
        raise RuntimeError(f"add_b200: {what} must be a CUDA tensor — this framework has no CPU fallback "
                           "(the CPU oracle lives in oracle/ and is test infrastructure only)")


def require_cuda_device(device: torch.device, what: str) -> None:
    if device.type != "cuda":
        raise RuntimeError(f"{what} needs the model on a CUDA device (add_b200 has no CPU fallback)")


class View:
    """Channel-slice view of a dense NHWC buffer (torch tensor of shape [N,H,W,Ctot])."""
    __slots__ = ("buf", "c_off", "c", "relud")

    def __init__(self, buf: torch.Tensor, c_off: int = 0, c: Optional[int] = None, relud: bool = False):
        assert buf.dim() == 4 and buf.is_contiguous()
        self.buf = buf
        self.c_off = c_off
        self.c = buf.shape[3] - c_off if c is None else c
        # relud: the buffer holds relu(value) because its producer stored it with RELU_OUT and every reader starts with
        # ReLU; convs reading it skip their ReLU-on-load pass (exact: relu is idempotent), raw readers must not see it
        self.relud = relud
        assert 0 <= c_off and c_off + self.c <= buf.shape[3]

    n = property(lambda s: s.buf.shape[0])
    h = property(lambda s: s.buf.shape[1])
    w = property(lambda s: s.buf.shape[2])
    dtype = property(lambda s: s.buf.dtype)

    def slice(self, off: int, c: int) -> "View":
        return View(self.buf, self.c_off + off, c, self.relud)

    def desc(self) -> AddTensor:
        b = self.buf
        return AddTensor(b.data_ptr() + self.c_off * b.element_size(), b.shape[0], b.shape[1], b.shape[2],
                         self.c, b.shape[3], _DTYPES[b.dtype])

    def nchw(self) -> torch.Tensor:
        """Logical NCHW tensor (channels_last strides) aliasing this view — zero copy."""
        t = self.buf[..., self.c_off:self.c_off + self.c]
        return t.permute(0, 3, 1, 2)


class ConvWeights:
    """A conv (+ folded eval-mode BN) ready for the kernels: fp32 [kh][kw][Cin][Cout] with the BN
    scale folded in, fp32 bias (w' = w·γ/σ, b' = β − μ·γ/σ), and, lazily, the
    bf16 UMMA-packed image for the tcgen05 path.

    The fold runs ON THE HOST (one-off preparation of a few MB of parameters): the parameters are copied down once,
    folded / permuted / padded in CPU fp32, and the finished images go up with plain memcpys — so no ATen device kernel
    is ever launched on behalf of this library and every kernel a profiler sees on the device is one of ours.
    `w_h` / `bias_h` are the host masters (derived weights — channel slices, merged FactorizedReduce taps, packed UMMA
    images — are built from them), `w` / `bias` the device copies the kernels read."""

    def __init__(self, weight: torch.Tensor, bn=None, bias: Optional[torch.Tensor] = None,
                 cin_pad: Optional[int] = None):
        dev = weight.device
        w = weight.detach().cpu().float()
        co, ci, kh, kw = w.shape
        if bn is not None:
            scale, shift = bn_scale_shift(bn)
            w = w * scale.view(-1, 1, 1, 1)
            b = shift if bias is None else shift + bias.detach().cpu().float() * scale
        else:
            b = bias.detach().cpu().float() if bias is not None else None
        w = w.permute(2, 3, 1, 0).contiguous()  # [kh][kw][Cin][Cout]
        if cin_pad is not None and cin_pad > ci:
            wp = torch.zeros(kh, kw, cin_pad, co, dtype=w.dtype)
            wp[:, :, :ci] = w
            w = wp
        self._finish(w, b, dev)

    def _finish(self, w_host: torch.Tensor, bias_host: Optional[torch.Tensor], dev) -> None:
        self.w_h = w_host.contiguous()
        self.bias_h = bias_host.contiguous() if bias_host is not None else None
        self.w = self.w_h.to(dev)
        self.bias = self.bias_h.to(dev) if self.bias_h is not None else None
        self.kh, self.kw, self.cin, self.cout = self.w_h.shape
        self._packed: Optional[torch.Tensor] = None

    @classmethod
    def from_folded(cls, w_hwio: torch.Tensor, bias: Optional[torch.Tensor], device=None) -> "ConvWeights":
        """From already folded HOST fp32 [kh][kw][Cin][Cout] weights (e.g. an input-channel slice of another conv's
        `w_h`); `device` = where the kernels will read them."""
        assert w_hwio.device.type == "cpu" and (bias is None or bias.device.type == "cpu"), "from_folded takes host masters"
        self = cls.__new__(cls)
        self._finish(w_hwio, bias, device if device is not None else w_hwio.device)
        return self

    def cout_slice(self, c0: int, g: int) -> "ConvWeights":
        """The conv restricted to output channels [c0, c0 + g) (cached)."""
        cache = self.__dict__.setdefault("_cout_slices", {})
        if (c0, g) not in cache:
            cs = ConvWeights.from_folded(self.w_h[..., c0:c0 + g].contiguous(),
                                         None if self.bias_h is None else self.bias_h[c0:c0 + g].contiguous(),
                                         self.w.device)
            if self.bias is not None:
                # `bias` is a public attribute (callers may have replaced the device vector after construction): the
                # slice reads the CURRENT device bias — a view, no copy kernel; c0 is a multiple of 256 -> 16-byte aligned
                cs.bias = self.bias[c0:c0 + g]
            cache[(c0, g)] = cs
        return cache[(c0, g)]

    def packed_tc(self) -> torch.Tensor:
        if self._packed is None:
            nbytes = lib.add_conv2d_tc_packed_bytes(self.cin, self.cout, self.kh, self.kw)
            if nbytes < 0:
                check(int(nbytes), "conv2d_tc_packed_bytes")
            host = torch.empty(nbytes, dtype=torch.uint8)
            check(lib.add_conv2d_tc_pack(self.w_h.data_ptr(), self.cin, self.cout, self.kh, self.kw,
                                         host.data_ptr()), "conv2d_tc_pack")
            self._packed = host.to(self.w.device)
        return self._packed


def pack_stem_tc(cw: "ConvWeights") -> torch.Tensor:
    """bf16 [64][64] UMMA image of a folded 3->64 3x3 conv for add_stem_conv3x3s2_nchw_fwd (device tensor)."""
    w_oihw = cw.w_h[:, :, :3, :].permute(3, 2, 0, 1).contiguous()        # [co][ci][ky][kx]
    assert tuple(w_oihw.shape) == (64, 3, 3, 3)
    host = torch.empty(int(lib.add_stem_tc_packed_bytes()), dtype=torch.uint8)
    check(lib.add_stem_tc_pack(w_oihw.data_ptr(), host.data_ptr()), "stem_tc_pack")
    return host.to(cw.w.device)


def bn_scale_shift(bn) -> Tuple[torch.Tensor, torch.Tensor]:
    """Eval-mode BatchNorm as y = x*scale + shift (HOST fp32 tensors)."""
    var = bn.running_var.detach().cpu().float()
    mean = bn.running_mean.detach().cpu().float()
    inv = torch.rsqrt(var + bn.eps)
    gamma = bn.weight.detach().cpu().float() if bn.weight is not None else torch.ones_like(var)
    beta = bn.bias.detach().cpu().float() if bn.bias is not None else torch.zeros_like(var)
    scale = gamma * inv
    return scale, beta - mean * scale


def host_to(t: torch.Tensor, dev) -> torch.Tensor:
    """A parameter prepared on the host (permuted / sliced, fp32 contiguous) -> device copy (a memcpy, no kernel)."""
    return t.contiguous().to(dev)


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# Which convs go to the tcgen05 path.  Set by conv_tc availability probing (see tc_available()).
_TC_STATE = {"enabled": True, "probed": None}


def set_tc_halo_mode(mode: int) -> None:
    """0 = per-tap A tiles, 1 = halo-resident A (default), 2 = halo-resident with descriptor base_offset."""
    check(lib.add_conv2d_tc_set_halo_mode(int(mode)), "conv2d_tc_set_halo_mode")
    bump_generation()


import os as _os
if _os.environ.get("ADD_PDL"):
    check(lib.add_set_pdl(int(_os.environ["ADD_PDL"])), "set_pdl")
if _os.environ.get("ADD_GRID_PCT"):
    check(lib.add_set_persistent_grid_pct(int(_os.environ["ADD_GRID_PCT"])), "set_persistent_grid_pct")
if _os.environ.get("ADD_TC_HALO_MODE"):
    set_tc_halo_mode(int(_os.environ["ADD_TC_HALO_MODE"]))
if _os.environ.get("ADD_BILINEAR_MODE"):
    check(lib.add_bilinear_set_mode(int(_os.environ["ADD_BILINEAR_MODE"])), "bilinear_set_mode")
if _os.environ.get("ADD_SEPCONV_MODE"):
    check(lib.add_sepconv_tc_set_mode(int(_os.environ["ADD_SEPCONV_MODE"])), "sepconv_tc_set_mode")


_GRAPH_STREAMS = {"n": int(_os.environ.get("ADD_GRAPH_STREAMS", "2"))}   # r3: 2 streams 822 img/s, 3: 808, 4: 800, 1: 779


def graph_streams() -> int:
    return _GRAPH_STREAMS["n"]


def set_graph_streams(n: int) -> None:
    """How many streams a captured plan's launch DAG is spread over (1 = one serial chain)."""
    _GRAPH_STREAMS["n"] = max(1, int(n))
    bump_generation()


def tc_available() -> bool:
    if _TC_STATE["probed"] is None:
        _TC_STATE["probed"] = lib.add_conv2d_tc_packed_bytes(64, 64, 1, 1) > 0
    return bool(_TC_STATE["probed"]) and _TC_STATE["enabled"]


def set_tc_enabled(flag: bool) -> None:
    _TC_STATE["enabled"] = bool(flag)
    bump_generation()


def _res(obj) -> Optional[tuple]:
    """Memory footprint of a launch operand for dependency tracking: (buffer base pointer, channel lo, channel hi).
    A View is a channel range of its buffer; a raw tensor covers its whole allocation."""
    if obj is None:
        return None
    if isinstance(obj, View):
        return (obj.buf.data_ptr(), obj.c_off, obj.c_off + obj.c)
    return (obj.data_ptr(), 0, 1 << 30)


def _overlap(a: tuple, b: tuple) -> bool:
    return a[0] == b[0] and a[1] < b[2] and b[1] < a[2]


class Builder:
    """Emits kernel launches.  record=False: launch immediately on the current stream.
    record=True: append to a launch list that `Plan` replays."""

    def __init__(self, device: torch.device, dtype: torch.dtype, record: bool = False):
        self.device = device
        self.dtype = dtype
        self.record = record
        self.launches: List[Tuple[Callable, tuple, str, dict]] = []
        self.keep: List[object] = []
        self._pool: Dict[tuple, List[torch.Tensor]] = {}
        self._resized: Dict[tuple, View] = {}
        self.bytes_allocated = 0

    # ---- memory ---------------------------------------------------------------------------
    def alloc(self, n: int, h: int, w: int, c: int, dtype: Optional[torch.dtype] = None) -> View:
        t = torch.empty((n, h, w, c), device=self.device, dtype=dtype or self.dtype)
        self.bytes_allocated += t.numel() * t.element_size()
        self.keep.append(t)
        return View(t)

    def scratch(self, n: int, h: int, w: int, c: int, dtype: Optional[torch.dtype] = None) -> View:
        if self.record:
            # recorded plans never recycle temporaries: a reused buffer is a false (write-after-read) dependency
            # that would serialise otherwise independent launches of the captured graph; HBM is not the constraint
            return self.alloc(n, h, w, c, dtype)
        key = (n, h, w, c, dtype or self.dtype)
        pool = self._pool.setdefault(key, [])
        if pool:
            return View(pool.pop())
        return self.alloc(n, h, w, c, dtype)

    def release(self, v: View) -> None:
        if self.record:
            return
        b = v.buf
        self._pool.setdefault((b.shape[0], b.shape[1], b.shape[2], b.shape[3], b.dtype), []).append(b)

    def raw(self, shape: Sequence[int], dtype: torch.dtype, zero: bool = False) -> torch.Tensor:
        t = (torch.zeros if zero else torch.empty)(tuple(shape), device=self.device, dtype=dtype)
        self.bytes_allocated += t.numel() * t.element_size()
        self.keep.append(t)
        return t

    # ---- launch plumbing ------------------------------------------------------------------
    def _emit(self, fn, args: tuple, tag: str, meta: Optional[dict] = None, reads=(), writes=()) -> None:
        """meta = {"kernel": name, "flops": algorithmic FLOPs, "bytes": algorithmic HBM bytes} of this
        launch (SURVEY §8d formulas) — what bench.py's roofline is computed from.  reads / writes: the Views /
        tensors the launch touches, from which `Plan` derives the launch DAG for multi-stream graph capture."""
        if self.record:
            meta = dict(meta or {"kernel": tag, "flops": 0, "bytes": 0})
            meta["reads"] = [r for r in map(_res, reads) if r is not None]
            meta["writes"] = [r for r in map(_res, writes) if r is not None]
            self.launches.append((fn, args, tag, meta))
        else:
            s = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            check(fn(*args, s), tag)

    def _d(self, v: View):
        d = v.desc()
        self.keep.append(d)
        return ctypes.byref(d)

    # ---- ops --------------------------------------------------------------------------------
    def conv(self, x: View, y: View, cw: ConvWeights, stride: int = 1, pad: int = 0, dil: int = 1,
             flags: int = 0, tag: str = "conv", image_bias: Optional[torch.Tensor] = None) -> None:
        """image_bias: fp32 [N, Cout] per-image bias replacing cw.bias (ASPP pool branch, see aspp_pool_bias)."""
        assert x.c == cw.cin and y.c == cw.cout, (x.c, cw.cin, y.c, cw.cout, tag)
        if (cw.cout > 256 and image_bias is None and x.dtype == torch.bfloat16 and tc_available() and stride <= 2
                and x.buf.shape[3] % 8 == 0 and x.c_off % 8 == 0):
            # the tcgen05 kernels hold at most 256 output channels (one TMEM accumulator row of fp32 columns): wider
            # convs (BASELINE config 5: C = 320 / 640 at F = 40 / 80) run as one launch per 256-channel output group
            for c0 in range(0, cw.cout, 256):
                g = min(256, cw.cout - c0)
                self.conv(x, y.slice(c0, g), cw.cout_slice(c0, g), stride, pad, dil, flags, tag)
            return
        self.keep.append(cw)
        if x.relud:
            assert flags & RELU_IN, f"{tag}: a post-ReLU buffer read by a conv that does not start with ReLU"
            flags &= ~RELU_IN
        bias_ptr, bias_stride = _ptr(cw.bias), 0
        extra_reads = ()
        if image_bias is not None:
            assert image_bias.dtype == torch.float32 and tuple(image_bias.shape) == (x.n, cw.cout) and image_bias.is_contiguous()
            bias_ptr, bias_stride, extra_reads = image_bias.data_ptr(), cw.cout, (image_bias,)
        # tcgen05 path: bf16 NHWC input whose base/stride are 16-byte multiples (TMA), Cout <= 256
        use_tc = (x.dtype == torch.bfloat16 and tc_available() and cw.cout <= 256 and stride <= 2
                  and x.buf.shape[3] % 8 == 0 and x.c_off % 8 == 0
                  and ((y.dtype == torch.bfloat16 and y.buf.shape[3] % 8 == 0 and y.c_off % 8 == 0)
                       or (y.dtype == torch.float32 and y.buf.shape[3] % 4 == 0 and y.c_off % 4 == 0)))
        p_out = y.n * y.h * y.w
        # strided 1x1 convs only touch the sampled lattice (SURVEY §8d FactorizedReduce row)
        p_in = p_out if (cw.kh == 1 and cw.kw == 1) else x.n * x.h * x.w
        ex, ey = x.buf.element_size(), y.buf.element_size()
        meta = dict(flops=2 * p_out * cw.cin * cw.cout * cw.kh * cw.kw,
                    bytes=p_in * cw.cin * ex + p_out * cw.cout * ey * (2 if flags & ACCUMULATE else 1)
                    + cw.cin * cw.cout * cw.kh * cw.kw * (2 if use_tc else 4))
        if use_tc:
            self._emit(lib.add_conv2d_tc_fwd,
                       (self._d(x), self._d(y), cw.packed_tc().data_ptr(), bias_ptr, bias_stride, cw.kh, cw.kw,
                        stride, pad, dil, flags), tag + ":tc", dict(kernel="conv2d_tc", **meta),
                       reads=((x, y) if flags & ACCUMULATE else (x,)) + extra_reads, writes=(y,))
        else:
            self._emit(lib.add_conv2d_fwd,
                       (self._d(x), self._d(y), cw.w.data_ptr(), bias_ptr, bias_stride, cw.kh, cw.kw,
                        stride, pad, dil, flags), tag, dict(kernel="conv2d_ffma", **meta),
                       reads=((x, y) if flags & ACCUMULATE else (x,)) + extra_reads, writes=(y,))

    def stem_nchw(self, src: torch.Tensor, y: View, w_packed: torch.Tensor, bias: torch.Tensor, flags: int,
                  tag: str = "stem0") -> None:
        """NCHW fp32 image -> 3x3 s2 conv 3->64 (+BN, ReLU) -> NHWC bf16, one launch (stem_tc.cu)."""
        n, c, h, w = src.shape
        assert c == 3 and src.dtype == torch.float32 and src.is_contiguous() and y.dtype == torch.bfloat16
        self.keep.extend((src, w_packed, bias))
        p_out = y.n * y.h * y.w
        self._emit(lib.add_stem_conv3x3s2_nchw_fwd,
                   (src.data_ptr(), n, h, w, self._d(y), w_packed.data_ptr(), bias.data_ptr(), flags), tag + ":tc",
                   dict(kernel="stem_conv_tc", flops=2 * p_out * 27 * y.c, bytes=src.numel() * 4 + p_out * y.c * 2 + 64 * 64 * 2),
                   reads=(src,), writes=(y,))

    def sepconv_half(self, x: View, y: View, w_dw: torch.Tensor, pw: ConvWeights, k: int, flags: int,
                     tag: str = "sephalf") -> None:
        self.keep.extend((w_dw, pw))
        if x.relud and (flags & RELU_IN):
            flags &= ~RELU_IN
        if x.c > 256 and x.dtype == torch.bfloat16 and tc_available() and x.c % 8 == 0:
            # wider than the fused tensor-core SepConv kernel takes (its pointwise GEMM has K = C <= 256; BASELINE
            # config 5: C = 320 / 640): stand-alone depthwise (bf16 result = the fused kernel's rounding point), then the
            # pointwise 1x1 + BN as a tcgen05 conv (split into 256-channel output groups by conv())
            mid = self.scratch(x.n, x.h, x.w, x.c)
            e = x.buf.element_size()
            self._emit(lib.add_depthwise_fwd, (self._d(x), self._d(mid), w_dw.data_ptr(), k, flags & RELU_IN), tag + ".dw",
                       dict(kernel="depthwise", flops=2 * x.n * x.h * x.w * x.c * k * k, bytes=2 * x.n * x.h * x.w * x.c * e),
                       reads=(x,), writes=(mid,))
            self.conv(mid, y, pw, 1, 0, 1, flags & ~RELU_IN, tag + ".pw")
            self.release(mid)
            return
        p = x.n * x.h * x.w
        ex = x.buf.element_size()
        # tensor-core path: bf16 NHWC input (TMA halo box), pointwise GEMM on tcgen05
        use_tc = (x.dtype == torch.bfloat16 and tc_available() and x.c % 8 == 0 and x.c <= 256 and y.c <= 256
                  and x.buf.shape[3] % 8 == 0 and x.c_off % 8 == 0
                  and ((y.dtype == torch.bfloat16 and y.buf.shape[3] % 8 == 0 and y.c_off % 8 == 0)
                       or (y.dtype == torch.float32 and y.buf.shape[3] % 4 == 0 and y.c_off % 4 == 0)))
        meta = dict(flops=2 * p * (x.c * k * k + x.c * y.c),
                    bytes=p * x.c * ex + p * y.c * y.buf.element_size() * (2 if flags & ACCUMULATE else 1)
                    + 4 * x.c * k * k + (2 if use_tc else 4) * x.c * y.c)
        if use_tc:
            self._emit(lib.add_sepconv_half_tc_fwd,
                       (self._d(x), self._d(y), w_dw.data_ptr(), pw.packed_tc().data_ptr(), _ptr(pw.bias), k, flags),
                       tag + ":tc", dict(kernel="sepconv_half_tc", **meta),
                       reads=(x, y) if flags & ACCUMULATE else (x,), writes=(y,))
        else:
            self._emit(lib.add_sepconv_half_fwd,
                       (self._d(x), self._d(y), w_dw.data_ptr(), pw.w.data_ptr(), _ptr(pw.bias), k, flags), tag,
                       dict(kernel="sepconv_half", **meta), reads=(x, y) if flags & ACCUMULATE else (x,), writes=(y,))

    def pool3x3(self, x: View, y: View, mode: int, stride: int, flags: int = 0, tag: str = "pool3x3") -> None:
        """mode 0 = avg_pool_3x3 (count_include_pad=False), 1 = max_pool_3x3 (operations.py:9-10)."""
        assert not x.relud, f"{tag}: raw read of a buffer that was stored post-ReLU"
        e = x.buf.element_size()
        meta = dict(kernel="pool3x3", flops=9 * y.n * y.h * y.w * y.c,
                    bytes=(x.n * x.h * x.w + y.n * y.h * y.w * (2 if flags & ACCUMULATE else 1)) * x.c * e)
        self._emit(lib.add_pool3x3_fwd, (self._d(x), self._d(y), mode, stride, flags), tag, meta,
                   reads=(x, y) if flags & ACCUMULATE else (x,), writes=(y,))

    def scale(self, x: View, y: View, scale: float, stride: int = 1, flags: int = 0, tag: str = "scale") -> None:
        """y (+)= scale * x[::stride, ::stride]: skip_connect (1.0) / none (0.0) (operations.py:8,11)."""
        assert not x.relud, f"{tag}: raw read of a buffer that was stored post-ReLU"
        e = x.buf.element_size()
        meta = dict(kernel="scale", flops=y.n * y.h * y.w * y.c,
                    bytes=y.n * y.h * y.w * y.c * e * (3 if flags & ACCUMULATE else 2))
        self._emit(lib.add_scale_fwd, (self._d(x), self._d(y), float(scale), stride, flags), tag, meta,
                   reads=(x, y) if flags & ACCUMULATE else (x,), writes=(y,))

    def bilinear(self, x: View, y: View, flags: int = 0, tag: str = "bilinear") -> None:
        assert not x.relud or (flags & RELU_IN), f"{tag}: raw read of a buffer that was stored post-ReLU"
        meta = dict(kernel="bilinear", flops=8 * y.n * y.h * y.w * y.c,
                    bytes=(x.n * x.h * x.w * x.buf.element_size() + y.n * y.h * y.w * y.buf.element_size()) * x.c)
        self._emit(lib.add_bilinear_fwd, (self._d(x), self._d(y), flags), tag, meta, reads=(x,), writes=(y,))

    def resized(self, x: View, h: int, w: int, tag: str = "bilinear") -> View:
        """Bilinear resize of a write-once tensor, computed once per (source, size) and shared by every later
        reader (the dense features of early cells are resized to the same few sizes by many cells, ADD.py:88-91)."""
        key = (x.buf.data_ptr(), x.c_off, x.c, h, w)
        v = self._resized.get(key)
        if v is None:
            v = self.alloc(x.n, h, w, x.c, x.dtype)
            self.bilinear(x, v, 0, tag)
            self._resized[key] = v
        return v

    def gather_images(self, src: torch.Tensor, dst: torch.Tensor, idx: torch.Tensor, tag: str = "gather_images") -> None:
        """dst[j] = src[idx[j]] over whole per-image slabs (dim 0 = image)."""
        assert src.shape[1:] == dst.shape[1:] and src.dtype == dst.dtype and idx.dtype == torch.int32
        per = src[0].numel() * src.element_size()
        self.keep.extend((src, dst, idx))
        self._emit(lib.add_gather_images, (src.data_ptr(), dst.data_ptr(), idx.data_ptr(), dst.shape[0], per), tag,
                   dict(kernel="gather_images", flops=0, bytes=2 * per * dst.shape[0]), reads=(src, idx), writes=(dst,))

    def gather_view(self, src: View, dst: View, idx: torch.Tensor, tag: str = "gather_view") -> None:
        """dst image j = src image idx[j], only the views' channels (both may be slices of wider buffers)."""
        assert (src.h, src.w, src.c, src.dtype) == (dst.h, dst.w, dst.c, dst.dtype) and idx.dtype == torch.int32
        self.keep.append(idx)
        e = src.buf.element_size()
        self._emit(lib.add_gather_images_view, (self._d(src), self._d(dst), idx.data_ptr()), tag,
                   dict(kernel="gather_images", flops=0, bytes=2 * dst.n * dst.h * dst.w * dst.c * e), reads=(src, idx), writes=(dst,))

    def gap(self, x: View, out: torch.Tensor, flags: int = 0, tag: str = "gap") -> None:
        nbytes = lib.add_global_avgpool_workspace_bytes(x.n, x.h, x.w, x.c)
        ws = self.raw((max(int(nbytes), 16),), torch.uint8)
        self._emit(lib.add_global_avgpool_fwd, (self._d(x), out.data_ptr(), flags, ws.data_ptr(), nbytes), tag,
                   dict(kernel="global_avgpool", flops=x.n * x.h * x.w * x.c, bytes=x.n * x.h * x.w * x.c * x.buf.element_size()),
                   reads=(x,), writes=(out, ws))

    def aspp_pool_bias(self, pooled: torch.Tensor, cw5: ConvWeights, w_out_pool: torch.Tensor,
                       b_out: Optional[torch.Tensor], out: torch.Tensor, tag: str = "aspp_pool_bias") -> None:
        """out[n][co] = b_out[co] + sum_d w_out_pool[d][co] * relu(b5[d] + sum_ci w5[ci][d] * pooled[n][ci])."""
        n, cin = pooled.shape[0], cw5.cin
        depth, cout = cw5.cout, out.shape[1]
        assert cw5.kh == 1 and tuple(w_out_pool.shape) == (depth, cout) and tuple(out.shape) == (n, cout)
        self.keep.extend((cw5, w_out_pool, b_out))
        self._emit(lib.add_aspp_pool_bias_fwd,
                   (pooled.data_ptr(), n, cin, cw5.w.data_ptr(), _ptr(cw5.bias), depth, w_out_pool.data_ptr(), _ptr(b_out),
                    cout, out.data_ptr()), tag,
                   dict(kernel="aspp_pool_bias", flops=2 * n * (cin * depth + depth * cout), bytes=4 * (cin * depth + depth * cout)),
                   reads=(pooled,), writes=(out,))

    def nchw_to_nhwc(self, src: torch.Tensor, c_src: int, y: View, tag: str = "nchw2nhwc") -> None:
        self.keep.append(src)
        self._emit(lib.add_nchw_to_nhwc, (src.data_ptr(), c_src, self._d(y)), tag,
                   dict(kernel="nchw_to_nhwc", flops=0, bytes=src.numel() * 4 + y.n * y.h * y.w * y.c * y.buf.element_size()),
                   reads=(src,), writes=(y,))

    def nhwc_to_nchw(self, x: View, dst: torch.Tensor, tag: str = "nhwc2nchw") -> None:
        self._emit(lib.add_nhwc_to_nchw, (self._d(x), dst.data_ptr()), tag,
                   dict(kernel="nhwc_to_nchw", flops=0, bytes=dst.numel() * 4 + x.n * x.h * x.w * x.c * x.buf.element_size()),
                   reads=(x,), writes=(dst,))

    def upsample_logits(self, x: View, dst: torch.Tensor, H: int, W: int, tag: str = "upsample_logits") -> None:
        self._emit(lib.add_upsample_logits_nchw, (self._d(x), dst.data_ptr(), H, W), tag,
                   dict(kernel="upsample_logits_nchw", flops=8 * dst.numel(), bytes=4 * (x.n * x.h * x.w * x.c + dst.numel())),
                   reads=(x,), writes=(dst,))

    def upsample_argmax(self, x: View, H: int, W: int, gt: Optional[torch.Tensor], pred: Optional[torch.Tensor],
                        cm: Optional[torch.Tensor], ent: Optional[torch.Tensor], tag: str = "upsample_argmax",
                        cm_rows: Optional[torch.Tensor] = None) -> None:
        """cm_rows: device int32 [n] — image j's confusion matrix goes to row cm_rows[j] of `cm` (scatter into the
        result of the original, uncompacted batch) instead of row j."""
        nbytes = lib.add_head_workspace_bytes(x.n, H, W, x.c)
        ws = self.raw((nbytes,), torch.uint8)
        u8 = gt is not None and gt.dtype == torch.uint8            # labels as the PNG bytes: 1 byte per pixel
        assert gt is None or gt.dtype in (torch.uint8, torch.int64)
        self._emit(lib.add_upsample_argmax_u8_fwd if u8 else lib.add_upsample_argmax_fwd,
                   (self._d(x), H, W, _ptr(gt), _ptr(pred), _ptr(cm), _ptr(cm_rows), _ptr(ent), ws.data_ptr(), nbytes), tag,
                   dict(kernel="upsample_argmax", flops=8 * x.n * H * W * x.c,
                        bytes=4 * x.n * x.h * x.w * x.c + x.n * H * W * ((1 if u8 else 8) * (gt is not None) + 8 * (pred is not None))),
                   reads=(x, gt, cm_rows), writes=(pred, cm, ent, ws))

    def edm_mlp(self, pooled: torch.Tensor, n: int, ws: Sequence[torch.Tensor], out: torch.Tensor,
                tag: str = "edm_mlp") -> None:
        self.keep.extend(ws)
        self._emit(lib.add_edm_mlp_fwd, (pooled.data_ptr(), n, *[t.data_ptr() for t in ws], out.data_ptr()), tag,
                   reads=(pooled,), writes=(out,))


class Plan:
    """A recorded launch list bound to its buffers; `run()` replays it on the current stream,
    `capture()` turns it into a CUDA graph (launch-bound inner loops: ~400 kernels per forward)."""

    def __init__(self, builder: Builder, start: int = 0, stop: Optional[int] = None):
        assert builder.record
        self.builder = builder
        self.launches = builder.launches[start:stop]
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.generation = generation()

    @property
    def n_launches(self) -> int:
        return len(self.launches)

    def run_eager(self) -> None:
        s = ctypes.c_void_p(torch.cuda.current_stream(self.builder.device).cuda_stream)
        for fn, args, tag, _ in self.launches:
            rc = fn(*args, s)
            if rc != 0:
                check(rc, tag)

    def profile(self) -> List[dict]:
        """Replay eagerly with a CUDA-event pair around every launch (on the launching stream) and
        return [{tag, kernel, flops, bytes, ms}] — per-launch device times for the roofline report."""
        stream = torch.cuda.current_stream(self.builder.device)
        s = ctypes.c_void_p(stream.cuda_stream)
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in self.launches]
        for (fn, args, tag, _), (e0, e1) in zip(self.launches, evs):
            e0.record(stream)
            rc = fn(*args, s)
            e1.record(stream)
            if rc != 0:
                check(rc, tag)
        stream.synchronize()
        return [dict(tag=tag, ms=e0.elapsed_time(e1), **{k: v for k, v in meta.items() if k not in ("reads", "writes")})
                for (_, _, tag, meta), (e0, e1) in zip(self.launches, evs)]

    # ---- launch DAG -------------------------------------------------------------------------
    def dependencies(self) -> List[List[int]]:
        """deps[i] = indices of the earlier launches launch i must wait for (RAW, WAW and WAR on channel ranges)."""
        deps: List[List[int]] = []
        hist: List[Tuple[tuple, int, bool]] = []            # (resource, launch index, is_write), newest last
        barrier = -1
        for i, (_, _, _, meta) in enumerate(self.launches):
            reads, writes = meta.get("reads", []), meta.get("writes", [])
            if not reads and not writes:                     # unknown footprint: full barrier
                deps.append(list(range(i)))
                barrier = i
                continue
            d = set() if barrier < 0 else {barrier}
            for r in reads:
                for res, j, is_w in hist:
                    if is_w and _overlap(r, res):
                        d.add(j)
            for w in writes:
                for res, j, _ in hist:
                    if _overlap(w, res):
                        d.add(j)
            deps.append(sorted(d))
            hist += [(r, i, False) for r in reads] + [(w, i, True) for w in writes]
        return deps

    def schedule(self, n_streams: int = 4) -> List[Tuple[int, List[int]]]:
        """Greedy list scheduling of the launch DAG onto `n_streams` streams, in program order.
        Returns [(stream index, [dependencies that live on OTHER streams])] per launch.  A launch continues the
        stream of its newest dependency when that dependency is still the tail of its stream (chains stay put);
        otherwise it goes to the stream whose tail is oldest (round-robin over idle lanes)."""
        deps = self.dependencies()
        stream_of: List[int] = []
        tail = [-1] * n_streams                               # last launch index per stream
        out = []
        for i, d in enumerate(deps):
            st = None
            if not self.launches[i][3].get("reads") and not self.launches[i][3].get("writes"):
                st = 0
            for j in sorted(d, reverse=True):
                if st is None and tail[stream_of[j]] == j:
                    st = stream_of[j]
                    break
            if st is None:
                st = min(range(n_streams), key=lambda k: tail[k])
            # a dependency on the same stream is implied by stream order; so is one covered transitively by a
            # later launch of that other stream that we already wait for — keep only the newest per other stream
            newest = {}
            for j in d:
                sj = stream_of[j]
                if sj != st:
                    newest[sj] = max(newest.get(sj, -1), j)
            out.append((st, sorted(newest.values())))
            stream_of.append(st)
            tail[st] = i
        return out

    def _run_streams(self, n_streams: int) -> None:
        """Issue the launches on several streams with event edges (inside a CUDA-graph capture: the graph gets
        the launch DAG's parallel branches instead of one serial chain)."""
        dev = self.builder.device
        main = torch.cuda.current_stream(dev)
        streams = [main] + [torch.cuda.Stream(dev) for _ in range(n_streams - 1)]
        for st in streams[1:]:
            st.wait_stream(main)
        sched = self.schedule(n_streams)
        needed = set(j for _, cross in sched for j in cross)
        events = {}
        for i, ((fn, args, tag, _), (si, cross)) in enumerate(zip(self.launches, sched)):
            st = streams[si]
            for j in cross:
                st.wait_event(events[j])
            rc = fn(*args, ctypes.c_void_p(st.cuda_stream))
            if rc != 0:
                check(rc, tag)
            if i in needed:
                ev = torch.cuda.Event()
                ev.record(st)
                events[i] = ev
        for st in streams[1:]:
            main.wait_stream(st)
        self._streams = streams                                # keep alive with the graph

    def capture(self, n_streams: Optional[int] = None) -> None:
        self.run_eager()  # warm-up: cudaFuncSetAttribute calls must not happen inside capture
        torch.cuda.current_stream(self.builder.device).synchronize()
        n_streams = graph_streams() if n_streams is None else n_streams
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            if n_streams > 1 and len(self.launches) > 1:
                self._run_streams(n_streams)
            else:
                self.run_eager()
        self.graph = g

    def run(self) -> None:
        if self.graph is not None:
            self.graph.replay()
        else:
            self.run_eager()


def as_nhwc_view(x: torch.Tensor, builder: Builder, dtype: torch.dtype, c_pad: Optional[int] = None) -> View:
    """Bring a logical-NCHW tensor (the reference's API layout) into an NHWC View of `dtype`.
    Zero copy when it already is channels_last in the right dtype with no padding needed."""
    require_cuda(x)
    n, c, h, w = x.shape
    cp = c_pad or c
    if x.dtype == dtype and cp == c and x.permute(0, 2, 3, 1).is_contiguous():
        return View(x.permute(0, 2, 3, 1))
    if x.dtype != torch.float32 or not x.is_contiguous():
        # dtype/stride normalisation of an API-edge tensor (plumbing, not the hot path)
        x = x.float().contiguous()
    y = builder.alloc(n, h, w, cp, dtype)
    builder.nchw_to_nhwc(x, c, y)
    return y
