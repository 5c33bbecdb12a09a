#!/usr/bin/env python
"""Per-op microbench on cuda:0 (BASELINE config 5 shapes): CUDA-event time of back-to-back launches, rotating
over several buffer sets so the working set is not trivially L2-resident.  Usage:
    python tools/microbench.py [filter-substring] [--reps N] [--once]      (--once: one launch per case, for ncu)"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import add_b200  # noqa: E402
from add_b200 import runtime as rt  # noqa: E402
from add_b200.runtime import Builder, ConvWeights, View, RELU_IN, RELU_OUT, ACCUMULATE  # noqa: E402

DEV = torch.device("cuda:0")
BN = torch.nn.BatchNorm2d


def cases():
    out = []
    for name, C, h, w in (("L1", 40, 128, 256), ("L2", 80, 64, 128), ("L3", 160, 32, 64), ("tiny", 40, 8, 16)):
        for op in ("sep_conv_3x3", "sep_conv_5x5", "dil_conv_3x3", "dil_conv_5x5"):
            out.append((f"{op}_{name}", "op", (op, C, 8, h, w)))
    out += [("pw_200to40_L1", "pw", (200, 40, 8, 128, 256)), ("pw_40to40_L1", "pw", (40, 40, 8, 128, 256)),
            ("pw_400to80_L2", "pw", (400, 80, 8, 64, 128)), ("pw_80to80_L2", "pw", (80, 80, 8, 64, 128)),
            ("pw_800to80_L2", "pw", (800, 80, 8, 64, 128)),
            ("dil5_L1_norelu", "conv", (40, 40, 5, 1, 4, 2, 8, 128, 256, 0)),
            ("dil5_L1_relu", "conv", (40, 40, 5, 1, 4, 2, 8, 128, 256, RELU_IN)),
            ("dil5_L1_half_rows", "conv", (40, 40, 5, 1, 4, 2, 8, 64, 256, RELU_IN)),
            ("dil5_L1_quarter", "conv", (40, 40, 5, 1, 4, 2, 8, 32, 256, RELU_IN)),
            ("stem1_64to64_3x3", "conv", (64, 64, 3, 1, 1, 1, 8, 512, 1024, RELU_OUT)),
            ("stem2_64to128_s2", "conv", (64, 128, 3, 2, 1, 1, 8, 512, 1024, RELU_IN)),
            ("aspp_400to256_d12", "conv", (400, 256, 3, 1, 12, 12, 4, 256, 512, RELU_IN | RELU_OUT)),
            ("aspp_400to256_d12_norelu", "conv", (400, 256, 3, 1, 12, 12, 4, 256, 512, RELU_OUT)),
            ("dec_304to256_3x3", "conv", (304, 256, 3, 1, 1, 1, 4, 128, 256, RELU_IN | RELU_OUT)),
            ("rowrate_dil3_c40", "conv", (40, 40, 3, 1, 2, 2, 8, 128, 256, 0)),
            ("rowrate_dil3_c64", "conv", (64, 64, 3, 1, 2, 2, 8, 128, 256, 0)),
            ("rowrate_dil3_c32", "conv", (32, 32, 3, 1, 2, 2, 8, 128, 256, 0)),
            ("rowrate_dil3_c40_d1", "conv", (40, 40, 3, 1, 1, 1, 8, 128, 256, 0)),
            ("rowrate_dil3_c40_d4", "conv", (40, 40, 3, 1, 4, 4, 8, 128, 256, 0)),
            ("rowrate_dil3_c40_d8", "conv", (40, 40, 3, 1, 8, 8, 8, 128, 256, 0)),
            ("rowrate_dil3_c64_d8", "conv", (64, 64, 3, 1, 8, 8, 8, 128, 256, 0)),
            ("rowrate_dil5_c40", "conv", (40, 40, 5, 1, 4, 2, 8, 128, 256, 0)),
            ("rowrate_dil5_c64", "conv", (64, 64, 5, 1, 4, 2, 8, 128, 256, 0)),
            ("bil_exit_63x127to256x512_c400", "bil", (400, 4, 63, 127, 256, 512)),
            ("bil_cellup_63x127to125x253_c400", "bil", (400, 4, 63, 127, 125, 253)),
            ("bil_dense_32x64to63x127_c160", "bil", (160, 4, 32, 64, 63, 127)),
            ("bil_down_256x512to128x256_c256", "bil", (256, 4, 256, 512, 128, 256))]
    return out


def sweep_cases():
    """BASELINE config 5: sep_conv / dil_conv 3x3 / 5x5 and ASPP_train at F = 20 / 40 / 80 across strides 4 / 8 / 16 / 32
    (C = F * stride / 4, spatial = 1024x2048 / stride, 8 images; ASPP input = 5C channels on 4 images)."""
    out = []
    for F in (20, 40, 80):
        for lvl, stride in enumerate((4, 8, 16, 32)):
            C, h, w = F * (1 << lvl), 1024 // stride, 2048 // stride
            for op in ("sep_conv_3x3", "sep_conv_5x5", "dil_conv_3x3", "dil_conv_5x5"):
                out.append((f"F{F}_s{stride}_{op}", "op", (op, C, 8, h, w)))
            out.append((f"F{F}_s{stride}_aspp_in{5 * C}", "aspp", (5 * C, 4, h, w)))
    return out


import os


def main():
    if os.environ.get("ADD_SEPCONV_MODE"):
        from add_b200._lib import lib
        assert lib.add_sepconv_tc_set_mode(int(os.environ["ADD_SEPCONV_MODE"])) == 0
    argv = sys.argv[1:]
    args = [a for i, a in enumerate(argv) if not a.startswith("--") and not (i > 0 and argv[i - 1] == "--reps")]
    filt = args[0] if args else ""
    reps = int(sys.argv[sys.argv.index("--reps") + 1]) if "--reps" in sys.argv else 20
    once = "--once" in sys.argv
    nbuf = 1 if once else 4
    torch.manual_seed(0)
    for name, kind, a in (sweep_cases() if "--sweep" in sys.argv else cases()):
        if filt not in name:
            continue
        b = Builder(DEV, torch.bfloat16, record=True)
        if kind == "aspp":
            cin, n, h, w = a
            try:
                m = add_b200.ASPP_train(cin, 256, BN, mult=1).eval().to(DEV)
                for _ in range(nbuf):
                    x = View(torch.randn(n, h, w, cin, device=DEV).to(torch.bfloat16)); b.keep.append(x.buf)
                    y = b.alloc(n, h, w, 256)
                    m.emit(b, x, y, 0)
            except Exception as e:      # shapes outside what the kernels take are reported, not hidden
                print(f"{name:22s} unsupported: {type(e).__name__}: {str(e)[:90]}")
                continue
            kind = "done"
        if kind == "done":
            pass
        elif kind == "op":
            op, C, n, h, w = a
            m = add_b200.OPS[op](C, 1, BN, 1e-5, 0.1, True).eval().to(DEV)
            try:
                for _ in range(nbuf):
                    x = View(torch.randn(n, h, w, C, device=DEV).to(torch.bfloat16)); b.keep.append(x.buf)
                    y = b.alloc(n, h, w, C)
                    m.emit(b, x, y, 0)
            except Exception as e:
                print(f"{name:22s} unsupported: {type(e).__name__}: {str(e)[:90]}")
                continue
        elif kind == "pw":
            cin, cout, n, h, w = a
            m = add_b200.ReLUConvBN(cin, cout, 1, 1, 0, BN).eval().to(DEV)
            for _ in range(nbuf):
                x = View(torch.randn(n, h, w, cin, device=DEV).to(torch.bfloat16)); b.keep.append(x.buf)
                y = b.alloc(n, h, w, cout)
                m.emit(b, x, y, 0)
        elif kind == "bil":
            C, n, h, w, ho, wo = a
            for _ in range(nbuf):
                x = View(torch.randn(n, h, w, C, device=DEV).to(torch.bfloat16)); b.keep.append(x.buf)
                y = b.alloc(n, ho, wo, C)
                b.bilinear(x, y, 0, name)
        else:
            cin, cout, k, stride, pad, dil, n, h, w, flags = a
            flags |= int(os.environ.get("ADD_MB_FLAGS", "0"), 0)
            cw = ConvWeights(torch.randn(cout, cin, k, k, device=DEV) / (cin * k * k) ** 0.5)
            cw.bias = torch.zeros(cout, device=DEV)
            ho, wo = (h + 2 * pad - dil * (k - 1) - 1) // stride + 1, (w + 2 * pad - dil * (k - 1) - 1) // stride + 1
            for _ in range(nbuf):
                x = View(torch.randn(n, h, w, cin, device=DEV).to(torch.bfloat16)); b.keep.append(x.buf)
                y = b.alloc(n, ho, wo, cout)
                b.conv(x, y, cw, stride, pad, dil, flags, name)
        plan = rt.Plan(b)
        try:
            plan.run_eager()
        except Exception as e:
            print(f"{name:22s} unsupported: {type(e).__name__}: {str(e)[:90]}")
            continue
        torch.cuda.synchronize()
        if once:
            continue
        plan.capture()
        for _ in range(3):
            plan.run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            plan.run()
        e1.record()
        torch.cuda.synchronize()
        n_launch = reps * len(plan.launches)
        us = 1e3 * e0.elapsed_time(e1) / (reps * nbuf)          # per op instance (a SepConv = 2 launches)
        fl = sum(l[3]["flops"] for l in plan.launches) / nbuf
        by = sum(l[3]["bytes"] for l in plan.launches) / nbuf
        kern = "+".join(sorted(set(l[3]["kernel"] for l in plan.launches)))
        print(f"{name:22s} {us:8.1f} us/op  {fl / us / 1e6:8.1f} TF/s  {by / us / 1e3:8.1f} GB/s  launches/op={len(plan.launches) // nbuf}  [{kern}]")


if __name__ == "__main__":
    main()
