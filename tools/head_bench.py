#!/usr/bin/env python
"""The two Evaluator kernels at BASELINE config-2 shapes, timed alone with CUDA events (rotating buffers larger than L2):
  * add_confusion_matrix            Evaluator._generate_matrix on int64 gt / pred  [8,1024,2048]   (16 B / pixel)
  * add_upsample_argmax_u8_fwd      x8 upsample + argmax + confusion matrix        4 images, logits [4,128,256,19] fp32
Prints achieved algorithmic GB/s against MEASURED_PEAKS.json.  Also the target of `ncu --set full -k regex:confusion_kernel|upsample_argmax`.
Usage: python tools/head_bench.py [reps] [--json out.json] [--variants]
--variants: time add_confusion_matrix with both histogram kernels (add_confusion_set_impl 0 / 1) and check them bit-identical."""
import ctypes
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import add_b200  # noqa: E402
from add_b200.runtime import Builder, View  # noqa: E402

_args = [a for a in sys.argv[1:] if not a.startswith("--")]
_json_out = sys.argv[sys.argv.index("--json") + 1] if "--json" in sys.argv else None
if _json_out in _args:
    _args.remove(_json_out)
reps = int(_args[0]) if _args else 20
RESULT = {}
dev = torch.device("cuda:0")
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
g = torch.Generator().manual_seed(1)
N, H, W = 8, 1024, 2048
sets = []
for i in range(3):                                       # 3 x 268 MB: every rep reads data that is not in the 126 MB L2
    gt = torch.randint(0, 19, (N, H, W), generator=g, dtype=torch.int64)
    gt[torch.rand(N, H, W, generator=g) < 0.1] = 255
    sets.append((gt.to(dev), torch.randint(0, 19, (N, H, W), generator=g, dtype=torch.int64).to(dev)))
ev = add_b200.Evaluator(19)


def timed(fn, n):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


by = N * H * W * 16
from add_b200._lib import lib as _lib  # noqa: E402
_variants = [(0, "per-warp match.any (default)")] + ([(1, "thread-private counters + wide finalize (opt-in)"), (2, "default histogram + wide finalize (opt-in)")]
                                                           if "--variants" in sys.argv else [])
_cms = []
for impl, label in _variants:
    assert _lib.add_confusion_set_impl(impl) == 0
    try:
        ms = timed(lambda i: ev._generate_matrix(*sets[i % 3]), reps)
        _cms.append(ev._generate_matrix(*sets[0]).cpu())
    finally:
        _lib.add_confusion_set_impl(0)
    print(f"confusion_matrix   {N}x{H}x{W} int64, {label}: {ms * 1e3:8.1f} us  {by / ms / 1e6:8.1f} GB/s  = {by / ms / 1e6 / peak:.3f} of measured HBM peak ({by / 1e6:.0f} MB algorithmic)")
    RESULT[f"confusion_matrix_impl{impl}"] = {"variant": label, "us": ms * 1e3, "gbs": by / ms / 1e6, "frac_of_hbm_peak": by / ms / 1e6 / peak,
                                              "algorithmic_bytes": by, "default": impl == 0}
if len(_cms) > 1:
    RESULT["confusion_variants_bit_identical"] = bool(all(torch.equal(_cms[0], c) for c in _cms[1:]))
    print("variants bit-identical:", RESULT["confusion_variants_bit_identical"])
if _json_out:
    Path(_json_out).write_text(json.dumps(RESULT))          # written early: the fused-head part below may be skipped by the caller's timeout

# fused head: low-res logits with spatially smooth classes (like a network's output) and a noisy variant (worst case)
n2 = 4
from add_b200.runtime import Plan  # noqa: E402
for name, smooth in (("smooth logits", True), ("iid-random logits", False)):
    bufs = []
    for i in range(3):
        if smooth:
            base = torch.randn(n2, 19, 16, 32, generator=g)
            lg = torch.nn.functional.interpolate(base, size=(128, 256), mode="bilinear") * 5 + torch.randn(n2, 19, 128, 256, generator=g) * 0.05
        else:
            lg = torch.randn(n2, 19, 128, 256, generator=g)
        buf = torch.zeros(n2, 128, 256, 20)
        buf[..., :19] = lg.permute(0, 2, 3, 1)
        gt8 = torch.randint(0, 19, (n2, H, W), generator=g, dtype=torch.int64).to(torch.uint8)
        cm = torch.zeros(n2, 19, 19, dtype=torch.int64, device=dev)
        bufs.append((View(buf.to(dev), 0, 19), gt8.to(dev), cm))
    plans = []
    for v_, g_, c_ in bufs:                      # recorded once: the workspace is allocated at record time, like in the network plans
        rb = Builder(dev, torch.float32, record=True)
        rb.upsample_argmax(v_, H, W, g_, None, c_, None)
        plans.append(Plan(rb))
    ms = timed(lambda i: plans[i % 3].run_eager(), reps)
    by = n2 * 128 * 256 * 19 * 4 + n2 * H * W
    print(f"upsample_argmax_cm {n2} images ({name}, uint8 labels): {ms * 1e3:8.1f} us  {by / ms / 1e6:8.1f} GB/s  = {by / ms / 1e6 / peak:.3f} of measured HBM peak ({by / 1e6:.1f} MB algorithmic)")
    RESULT[f"upsample_argmax_cm_{'smooth' if smooth else 'iid'}"] = {"us": ms * 1e3, "gbs": by / ms / 1e6, "algorithmic_bytes": by, "images": n2}
    # bit-exact against materialised logits -> argmax -> Evaluator
    v, gt8, cm = bufs[0]
    out = torch.empty(n2, 19, H, W, device=dev)
    Builder(dev, torch.float32, record=False).upsample_logits(v, out, H, W)
    for j in range(n2):
        want = add_b200.Evaluator(19)._generate_matrix(gt8[j].long(), out[j:j + 1].argmax(1)[0])
        assert torch.equal(want, cm[j]), (name, j)
if _json_out:
    Path(_json_out).write_text(json.dumps(RESULT))
print("ok")
