"""GPU: the tcgen05/TMEM implicit-GEMM convolution against the CUDA-core fp32 kernel on identical
bf16 inputs and bf16-representable weights, so both compute the same products with fp32
accumulation: the only difference is summation order.  Tolerance 2e-5 max-norm relative for fp32
outputs; 1 bf16 ulp (2^-8 relative) for bf16 outputs."""
import numpy as np
import pytest
import torch

import util
import add_b200
from add_b200 import runtime as rt
from add_b200.runtime import Builder, ConvWeights, View, RELU_IN, RELU_OUT, ACCUMULATE

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _bf16_exact(t):
    return t.to(torch.bfloat16).float()


def _run(x_buf, c_off, cin, w, bias, stride, pad, dil, flags, out_hw, out_dtype, y_ctot, y_off, tc, y_init):
    rt.set_tc_enabled(tc)
    b = Builder(DEV, torch.bfloat16)
    xv = View(x_buf, c_off, cin)
    ybuf = y_init.clone()
    yv = View(ybuf, y_off, w.shape[0])
    cw = ConvWeights(w)
    cw.bias = bias
    b.conv(xv, yv, cw, stride, pad, dil, flags)
    torch.cuda.synchronize()
    rt.set_tc_enabled(True)
    return ybuf


CASES = [
    # cin, cout, k, stride, pad, dil, H, W, flags
    (64, 64, 1, 1, 0, 1, 16, 128, 0),
    (40, 40, 1, 1, 0, 1, 13, 17, RELU_IN),
    (200, 40, 1, 1, 0, 1, 31, 64, RELU_IN),
    (400, 80, 1, 1, 0, 1, 9, 130, RELU_IN | RELU_OUT),
    (800, 80, 1, 1, 0, 1, 7, 33, RELU_IN),
    (256, 19, 1, 1, 0, 1, 16, 40, 0),
    (40, 40, 3, 1, 2, 2, 20, 37, RELU_IN),
    (40, 40, 5, 1, 4, 2, 20, 37, RELU_IN | ACCUMULATE),
    (80, 80, 5, 1, 4, 2, 11, 64, RELU_IN),
    (160, 160, 3, 1, 2, 2, 8, 16, RELU_IN | ACCUMULATE),
    (304, 256, 3, 1, 1, 1, 12, 24, RELU_IN | RELU_OUT),
    (400, 256, 3, 1, 6, 6, 16, 32, RELU_IN | RELU_OUT),
    (400, 256, 3, 1, 18, 18, 16, 32, RELU_OUT),
    (8, 64, 3, 2, 1, 1, 32, 64, RELU_OUT),
    (64, 128, 3, 2, 1, 1, 17, 33, RELU_IN),
    (128, 24, 1, 2, 0, 1, 12, 16, RELU_IN),
    (128, 24, 1, 2, -1, 1, 13, 9, RELU_IN),
    (400, 128, 3, 2, 1, 1, 16, 32, RELU_IN | RELU_OUT),
    (64, 64, 3, 1, 1, 1, 5, 300, RELU_OUT),
    # halo-resident path (stride 1, Wo > 64, Cout <= 160): shifted-descriptor taps
    (40, 40, 5, 1, 4, 2, 9, 253, RELU_IN | ACCUMULATE),
    (40, 40, 3, 1, 2, 2, 7, 127, RELU_IN),
    (80, 80, 5, 1, 4, 2, 6, 128, RELU_IN | RELU_OUT),
    (160, 160, 3, 1, 2, 2, 5, 100, RELU_IN | ACCUMULATE),
    (200, 48, 3, 1, 1, 1, 4, 65, 0),
    (64, 64, 3, 1, 1, 1, 3, 1024, RELU_OUT),
    (40, 40, 5, 1, 2, 1, 11, 256, RELU_IN),
    # more tiles than resident CTAs (persistent kernel: ring continues across tiles, TMEM double buffer, resident B)
    (40, 40, 1, 1, 0, 1, 160, 250, RELU_IN | ACCUMULATE),
    (200, 40, 1, 1, 0, 1, 150, 253, RELU_IN),
    (400, 80, 1, 1, 0, 1, 127, 200, RELU_IN | RELU_OUT),
    (64, 64, 3, 1, 1, 1, 130, 200, RELU_OUT),
    (64, 128, 3, 2, 1, 1, 257, 300, RELU_IN),
    # persistent halo kernel: more row tiles than SMs; resident weights (3x3 C<=64, 5x5 ring), two channel chunks
    (40, 40, 5, 1, 4, 2, 100, 253, RELU_IN | ACCUMULATE),
    (40, 40, 3, 1, 2, 2, 110, 250, 0),
    (64, 64, 3, 1, 1, 1, 90, 300, RELU_OUT),
    (80, 80, 3, 1, 2, 2, 120, 127, RELU_IN),
    (80, 80, 5, 1, 4, 2, 100, 128, RELU_IN | ACCUMULATE),
    # large enough (>= 4 x SMs units of two tiles) for the cluster-of-2 path: weight halves TMA-multicast to both CTAs
    (400, 256, 3, 1, 6, 6, 64, 1280, RELU_IN | RELU_OUT),
    (304, 256, 3, 1, 1, 1, 75, 1100, RELU_OUT),
    # row pairs (r, r + dil): odd heights leave unpaired / out-of-image rows
    (40, 40, 5, 1, 4, 2, 125, 253, RELU_IN),
    (40, 40, 3, 1, 2, 2, 63, 127, RELU_IN | ACCUMULATE),
    (64, 64, 3, 1, 1, 1, 131, 200, RELU_OUT),
    (80, 80, 5, 1, 4, 2, 63, 127, 0),
    # row quads (r, r + dil, r + 2 dil, r + 3 dil) for the streamed-weight 5x5s once the quads cover the SMs; per-row halo
    # barriers; heights that are not a multiple of 4 * dil; two channel chunks with a single accumulator set
    (40, 40, 5, 1, 4, 2, 131, 300, RELU_IN | ACCUMULATE),
    (80, 80, 5, 1, 4, 2, 250, 128, RELU_IN),
    (40, 40, 5, 1, 2, 1, 243, 256, 0),
    # more than 256 output channels (BASELINE config 5: C = 320 / 640): one tcgen05 launch per 256-channel output group
    (320, 320, 3, 1, 2, 2, 8, 16, RELU_IN),
    (640, 640, 5, 1, 4, 2, 6, 9, RELU_IN | ACCUMULATE),
    (1600, 256, 1, 1, 0, 1, 5, 12, RELU_IN | RELU_OUT),
]


@pytest.mark.parametrize("case", CASES, ids=[f"c{c[0]}-{c[1]}_k{c[2]}s{c[3]}p{c[4]}d{c[5]}_{c[6]}x{c[7]}_f{c[8]}" for c in CASES])
@pytest.mark.parametrize("persistent", [True, False], ids=["persistent", "tile_per_cta"])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_tc_matches_ffma(case, out_dtype, persistent):
    cin, cout, k, stride, pad, dil, H, W, flags = case
    rt.set_tc_halo_mode(1 if persistent else 17)        # bit 4: one tile per CTA instead of the persistent kernel
    g = torch.Generator().manual_seed(hash(case) % (2 ** 31))
    n = 2
    # input is a channel slice of a wider buffer (concat-slice reads)
    x_ctot, c_off = cin + 16, 8
    x_buf = torch.randn(n, H, W, x_ctot, generator=g).to(torch.bfloat16).to(DEV)
    w = _bf16_exact(torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).to(DEV)
    bias = torch.randn(cout, generator=g).to(DEV)
    if pad >= 0:
        ho = (H + 2 * pad - dil * (k - 1) - 1) // stride + 1
        wo = (W + 2 * pad - dil * (k - 1) - 1) // stride + 1
    else:
        ho, wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    step = 8 if out_dtype == torch.bfloat16 else 4
    y_ctot, y_off = (cout + step - 1) // step * step + 2 * step, step
    y_init = torch.randn(n, ho, wo, y_ctot, generator=g).to(out_dtype).to(DEV)
    ref = _run(x_buf, c_off, cin, w, bias, stride, pad, dil, flags, (ho, wo), out_dtype, y_ctot, y_off, False, y_init)
    try:
        got = _run(x_buf, c_off, cin, w, bias, stride, pad, dil, flags, (ho, wo), out_dtype, y_ctot, y_off, True, y_init)
    finally:
        rt.set_tc_halo_mode(1)
    # untouched channels stay untouched
    assert torch.equal(got[..., :y_off], y_init[..., :y_off])
    assert torch.equal(got[..., y_off + cout:], y_init[..., y_off + cout:])
    a, r = got[..., y_off:y_off + cout].float(), ref[..., y_off:y_off + cout].float()
    tol = 2e-5 if out_dtype == torch.float32 else 2 ** -7
    assert util.rel_err(a, r) < tol


def test_tc_is_the_bf16_conv_path():
    assert rt.tc_available()
    m, x = util.make_op_case("dil_conv_5x5_c40")
    m = m.to(DEV)
    b = Builder(DEV, torch.bfloat16, record=True)
    xv = rt.as_nhwc_view(x.to(DEV).to(torch.bfloat16).contiguous(memory_format=torch.channels_last), b, torch.bfloat16)
    yv = b.alloc(xv.n, xv.h, xv.w, 40)
    m.emit(b, xv, yv, 0)
    assert [l[3]["kernel"] for l in b.launches] == ["conv2d_tc"]


def test_network_bf16_vs_fp32():
    """Whole network in bf16 (tensor-core path) against the fp32 golden.  Stated per-exit tolerance
    (max-norm relative, BN-randomised weights): exit 1 (6 cells deep) 3e-2, last exit (12 cells deep)
    1e-1 — every activation is rounded to bf16 (2^-8) at each of ~50 sequential layers; argmax
    agreement >= 95 %."""
    NETS = np.load(util.ROOT / "tests/golden/nets.npz")
    cname = "searched-dense-C2"
    spec = util.NET_CASES[cname]
    net = util.make_net(spec).to(DEV)
    net.set_precision("bf16")
    for (h, w) in spec["sizes"]:
        x, _ = util.make_input(1, h, w)
        outs = net(x.to(DEV))
        for e, o in enumerate(outs):
            ref = torch.from_numpy(NETS[f"{cname}/{h}x{w}/forward/{e}"])
            err = util.rel_err(o, ref)
            agree = float((o.cpu().argmax(1) == ref.argmax(1)).float().mean())
            print(f"bf16 {h}x{w} exit {e}: rel_err={err:.3e} argmax_agree={agree:.4f}")
            assert err < (3e-2 if e < len(outs) - 1 else 1e-1)
            assert agree > 0.95


# ---- SepConv half on the tensor-core path (sepconv_tc.cu) vs a plain PyTorch fp32 reference ----------
SEP_CASES = [
    # C, Cout, k, H, W, flags
    (40, 40, 3, 13, 17, RELU_IN | RELU_OUT),
    (40, 40, 5, 13, 17, RELU_IN | RELU_OUT),
    (40, 40, 5, 8, 16, 0),
    (40, 40, 3, 21, 253, ACCUMULATE),
    (80, 80, 3, 9, 20, RELU_IN | RELU_OUT),
    (80, 80, 5, 17, 127, RELU_IN | ACCUMULATE),
    (160, 160, 5, 6, 7, RELU_IN),
    (160, 160, 3, 32, 64, ACCUMULATE | RELU_OUT),
    (24, 24, 3, 7, 9, RELU_IN),
    (64, 48, 5, 10, 33, RELU_IN | RELU_OUT),
    (200, 40, 3, 5, 18, 0),
    # more tiles than resident CTAs: the persistent kernel's multi-tile loop, ring wrap-around and phase flips
    (40, 40, 5, 100, 253, RELU_IN | ACCUMULATE),
    (40, 40, 3, 131, 250, RELU_IN | RELU_OUT),
    (80, 80, 3, 90, 127, RELU_IN | RELU_OUT),
    (80, 80, 5, 70, 130, ACCUMULATE),
    # many tiles per CTA: A / TMEM / halo rings wrapping several times
    (40, 40, 3, 260, 500, RELU_IN | ACCUMULATE),
    (40, 40, 5, 250, 400, RELU_OUT),
    (80, 80, 3, 200, 260, 0),
]


def _sep_reference(x, w_dw, w_pw, bias, k, flags, y_init):
    """fp32 PyTorch: relu? -> depthwise (fp32) -> round to bf16 (the kernel's A operand) -> 1x1 -> +bias."""
    import torch.nn.functional as F
    xin = x.float().permute(0, 3, 1, 2)
    if flags & RELU_IN:
        xin = F.relu(xin)
    C = xin.shape[1]
    d = F.conv2d(xin, w_dw.permute(2, 0, 1).unsqueeze(1), padding=k // 2, groups=C)
    d = d.to(torch.bfloat16).float()
    o = F.conv2d(d, w_pw.t().reshape(w_pw.shape[1], C, 1, 1)) + bias.view(1, -1, 1, 1)
    o = o.permute(0, 2, 3, 1)
    if flags & ACCUMULATE:
        o = o + y_init.float()
    if flags & RELU_OUT:
        o = F.relu(o)
    return o


@pytest.mark.parametrize("case", SEP_CASES, ids=[f"c{c[0]}-{c[1]}_k{c[2]}_{c[3]}x{c[4]}_f{c[5]}" for c in SEP_CASES])
@pytest.mark.parametrize("mode", [1, 0, 3], ids=["persistent", "tile_per_cta", "persistent_unmerged"])
@pytest.mark.parametrize("contig", [False, True], ids=["slice_in", "contig_in"])
@pytest.mark.parametrize("out_dtype", [torch.bfloat16, torch.float32])
def test_sepconv_half_tc_matches_torch(case, out_dtype, mode, contig):
    """Tolerance: fp32 output 2e-3, bf16 output 2^-7 (max-norm relative).  The depthwise result is
    rounded to bf16 before the pointwise GEMM in both; fp32 summation order may flip a rounding."""
    C, Cout, k, H, W, flags = case
    g = torch.Generator().manual_seed(hash(case) % (2 ** 31))
    n = 2
    x_ctot, c_off = (C, 0) if contig else (C + 16, 8)     # contiguous input -> merged {W*C} halo tensor map
    x_buf = torch.randn(n, H, W, x_ctot, generator=g).to(torch.bfloat16).to(DEV)
    w_dw = (torch.randn(k, k, C, generator=g) / k).to(DEV)
    w_pw = _bf16_exact(torch.randn(C, Cout, generator=g) / C ** 0.5).to(DEV)     # [Cin][Cout]
    bias = torch.randn(Cout, generator=g).to(DEV)
    step = 8 if out_dtype == torch.bfloat16 else 4
    y_ctot, y_off = (Cout + step - 1) // step * step + 2 * step, step
    y_init = torch.randn(n, H, W, y_ctot, generator=g).to(out_dtype).to(DEV)
    cw = ConvWeights(w_pw.t().reshape(Cout, C, 1, 1).contiguous())
    cw.bias = bias
    b = Builder(DEV, torch.bfloat16, record=True)
    ybuf = y_init.clone()
    b.sepconv_half(View(x_buf, c_off, C), View(ybuf, y_off, Cout), w_dw.contiguous(), cw, k, flags)
    assert [l[3]["kernel"] for l in b.launches] == ["sepconv_half_tc"]
    from add_b200._lib import lib as _lib
    assert _lib.add_sepconv_tc_set_mode(mode) == 0
    try:
        rt.Plan(b).run_eager()
        torch.cuda.synchronize()
    finally:
        _lib.add_sepconv_tc_set_mode(1)
    ref = _sep_reference(x_buf[..., c_off:c_off + C], w_dw, w_pw, bias, k, flags, y_init[..., y_off:y_off + Cout])
    assert torch.equal(ybuf[..., :y_off], y_init[..., :y_off])
    assert torch.equal(ybuf[..., y_off + Cout:], y_init[..., y_off + Cout:])
    tol = 2e-3 if out_dtype == torch.float32 else 2 ** -7
    assert util.rel_err(ybuf[..., y_off:y_off + Cout].float(), ref) < tol


# ---- fused stem0 (stem_tc.cu): NCHW fp32 image -> conv3x3 s2 + bias + ReLU -> NHWC bf16 -------------------
@pytest.mark.parametrize("k", [3, 5])
@pytest.mark.parametrize("C", [320, 640])
def test_sepconv_half_wider_than_the_fused_kernel(k, C):
    """C = 320 / 640 (BASELINE config 5, F = 40 / 80 at the deep strides): wider than the fused tensor-core SepConv
    kernel takes, so a half runs as a stand-alone depthwise (bf16 result, the fused kernel's rounding point) plus the
    pointwise 1x1 on the tcgen05 path in 256-channel output groups.  Same reference and tolerance as the fused kernel."""
    H, W, flags = 9, 14, RELU_IN | RELU_OUT
    g = torch.Generator().manual_seed(71 + k + C)
    x_buf = torch.randn(2, H, W, C, generator=g).to(torch.bfloat16).to(DEV)
    w_dw = (torch.randn(k, k, C, generator=g) / k).to(DEV)
    w_pw = _bf16_exact(torch.randn(C, C, generator=g) / C ** 0.5).to(DEV)
    bias = torch.randn(C, generator=g).to(DEV)
    cw = ConvWeights(w_pw.t().reshape(C, C, 1, 1).contiguous())
    cw.bias = bias
    b = Builder(DEV, torch.bfloat16, record=True)
    y = torch.zeros(2, H, W, C, dtype=torch.bfloat16, device=DEV)
    b.sepconv_half(View(x_buf, 0, C), View(y, 0, C), w_dw.contiguous(), cw, k, flags)
    kernels = [l[3]["kernel"] for l in b.launches]
    assert kernels[0] == "depthwise" and set(kernels[1:]) == {"conv2d_tc"} and len(kernels) == 1 + (C + 255) // 256
    rt.Plan(b).run_eager()
    torch.cuda.synchronize()
    ref = _sep_reference(x_buf, w_dw, w_pw, bias, k, flags, y)
    assert util.rel_err(y.float(), ref) < 2 ** -7


@pytest.mark.parametrize("hw", [(32, 64), (33, 65), (48, 300), (17, 513)])
def test_stem_tc_matches_torch(hw):
    """vs plain PyTorch fp32 conv on bf16-rounded image and weights (what the kernel multiplies); output is bf16:
    tolerance 2^-7 max-norm relative."""
    import torch.nn.functional as F
    H, W = hw
    g = torch.Generator().manual_seed(H * 1000 + W)
    x = torch.randn(2, 3, H, W, generator=g)
    w = _bf16_exact(torch.randn(64, 3, 3, 3, generator=g) / 27 ** 0.5)
    bias = torch.randn(64, generator=g)
    cw = ConvWeights(w.to(DEV))
    cw.bias = bias.to(DEV)
    packed = rt.pack_stem_tc(cw)
    b = Builder(DEV, torch.bfloat16)
    y = b.alloc(2, (H - 1) // 2 + 1, (W - 1) // 2 + 1, 64)
    y.buf.fill_(float("nan"))
    b.stem_nchw(x.to(DEV), y, packed, cw.bias, RELU_OUT)
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(_bf16_exact(x), w, bias, stride=2, padding=1)).permute(0, 2, 3, 1)
    assert util.rel_err(y.buf.float(), ref) < 2 ** -7


@pytest.mark.parametrize("tc", [True, False], ids=["tcgen05", "ffma"])
@pytest.mark.parametrize("hw", [(9, 40), (40, 300)])
def test_per_image_bias_matches_torch(tc, hw):
    """bias_image_stride (ASPP pool branch folded into the 1x1): image n gets its own bias vector.  vs plain PyTorch
    fp32 on bf16-representable operands; fp32 output, tolerance 2e-5 max-norm relative."""
    import torch.nn.functional as F
    H, W = hw
    g = torch.Generator().manual_seed(H * 31 + W)
    n, cin, cout = 3, 128, 64
    x = torch.randn(n, H, W, cin, generator=g).to(torch.bfloat16).to(DEV)
    w = _bf16_exact(torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5).to(DEV)
    bias_n = torch.randn(n, cout, generator=g).to(DEV)
    rt.set_tc_enabled(tc)
    try:
        b = Builder(DEV, torch.bfloat16)
        y = b.alloc(n, H, W, cout, torch.float32)
        b.conv(View(x), y, ConvWeights(w), 1, 0, 1, RELU_IN, image_bias=bias_n)
        torch.cuda.synchronize()
    finally:
        rt.set_tc_enabled(True)
    ref = F.conv2d(F.relu(x.float().permute(0, 3, 1, 2)), w) + bias_n.view(n, cout, 1, 1)
    assert util.rel_err(y.buf, ref.permute(0, 2, 3, 1)) < 2e-5
