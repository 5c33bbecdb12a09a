#!/usr/bin/env python
"""bench.py — ADD 1024x2048 inference images/sec on N B200s (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (libadd_b200, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the UNMODIFIED reference (baseline/_ref) on the host cores

One step = one pass of the hot path over one batch: BASELINE config 2 — searched-dense ADD (C=2,
F=20), 8 synthetic 3x1024x2048 images per GPU, bf16, EDM-gated early exit applied per image
(threshold at the batch median ⇒ 4 of 8 images exit early; reference/parity exit semantics), each
image's exit → argmax → int64 confusion matrix (what eval.py:195-221 does per image).
`value`: inputs already resident in HBM.  `e2e`: the public API call with HOST (pinned) buffers —
H2D of images+labels and D2H of the confusion matrices inside the timed region.
Multi-GPU: batch sharding, one process per GPU, no data-path collective (weak scaling).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "ADD 1024x2048 inference images/sec"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "torch_cuda"],
                    help="b200 = this repo; reference = the UNMODIFIED reference on the host cores (baseline/_ref; the oracle port only if nothing was staged); torch_cuda = informational "
                         "stock PyTorch/cuDNN comparator on cuda:0")
    ap.add_argument("--batch", type=int, default=8, help="images per GPU per step")
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=2048)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--exit-mode", default="reference", choices=["reference", "forward"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--forward-only", action="store_true",
                    help="informational: time ADD.evaluate (all exits, no gating) on the resident batch and exit")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="resident arm: one blocking dynamic_evaluate call per step instead of the software pipeline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-images", type=int, default=12, help="images in the bounded CPU-baseline sample")
    ap.add_argument("--label-dtype", default="uint8", choices=["uint8", "int64"],
                    help="dtype of the HOST labels fed to the e2e loop (uint8 = Cityscapes PNG depth, widened on the device)")
    ap.add_argument("--image-dtype", default="uint8", choices=["uint8", "fp32"],
                    help="dtype of the HOST images fed to the e2e loop: uint8 HWC (the PNG bytes; normalised on the device "
                         "with the reference loader's arithmetic) or the loader's normalised fp32 NCHW")
    ap.add_argument("--per-rank-threshold", action="store_true",
                    help="multi-GPU: every rank gates at ITS OWN batch median (always 50 %% early exits per rank, the round-1 "
                         "behaviour).  Default: ONE global threshold — rank 0's batch median, broadcast before the timed "
                         "region — as a deployment would use, so exit counts (and step times) differ per rank")
    ap.add_argument("--no-fp32-feed", action="store_true",
                    help="skip the second e2e measurement that feeds the reference loader's fp32 NCHW images + int64 labels")
    ap.add_argument("--no-candidates", action="store_true",
                    help="skip the child-process measurement of the opt-in candidate kernels (`candidates` in the JSON line)")
    ap.add_argument("--ref-budget-s", type=float, default=150.0,
                    help="--impl reference: wall-clock budget of the whole run; the step shrinks to a bounded sample of the batch to fit")
    ap.add_argument("--profile-out", default=None, help="write the per-kernel launch table (JSON) here")
    ap.add_argument("--profile-step", action="store_true",
                    help="for `ncu --profile-from-start off`: warm up, then run ONE step between cudaProfilerStart/Stop "
                         "and exit (prints no bench line)")
    return ap.parse_args()


def workload_name(a) -> dict:
    return {"workload": f"searched-dense ADD C=2 F=20, {a.batch}x3x{a.height}x{a.width} per GPU, "
                        f"EDM-gated early exit per image (threshold = batch median, {a.exit_mode} exit semantics), "
                        "argmax + confusion matrix per exit taken",
            "network": "searched-dense", "C": 2, "F": 20, "batch_per_gpu": a.batch,
            "height": a.height, "width": a.width, "exit_mode": a.exit_mode,
            "weights": "random init, seed 1 (kaiming conv, BN identity stats)",
            "l2": "inputs per step (%.0f MB) exceed the 126 MB L2" % (a.batch * 3 * a.height * a.width * 4 / 1e6)}


# --------------------------------------------------------------------------------------------------
# clocks sampler (B200_PROFILING.md recipe)
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    """`nvidia-smi -lms 20` in the background.  It is started BEFORE the warm-up (the tool takes a few hundred ms
    to produce its first line); `mark()` brackets the timed region and only samples that arrived inside it are
    reported (falling back to the samples under load since start if the region was too short to catch one)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines, self.marks = gpu_index, None, [], []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "10", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
            t0 = time.perf_counter()
            while not self.lines and time.perf_counter() - t0 < 3.0:     # nvidia-smi needs a few hundred ms to start
                time.sleep(0.01)
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def mark(self):
        self.marks.append(time.perf_counter())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def collect(lo, hi):
            sm, smax, reasons = [], [], set()
            for ts, ln in self.lines:
                if not (lo <= ts <= hi):
                    continue
                f = [t.strip() for t in ln.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); smax.append(float(f[2]))
                except ValueError:
                    continue
                for nm, val in zip(names, f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(nm)
            return sm, smax, reasons
        lo, hi = (self.marks + [0.0, float("inf")])[:2] if len(self.marks) >= 2 else (0.0, float("inf"))
        sm, smax, reasons = collect(lo, hi + 0.03)
        window = "timed region"
        if not sm:
            sm, smax, reasons = collect(0.0, float("inf"))
            window = "warm-up + timed region (timed region shorter than one sample)"
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "window": window, "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the reference's own CPU path on all host threads.
# Nothing here imports add_b200 (the product): the reference arm's process must not map libadd_b200.so.
#   kind "reference": the UNMODIFIED reference modules staged under the git-ignored baseline/_ref/
#                     (oracle/stage_reference.py) — modeling.ADD.ADD.dynamic_inference + utils.metrics.Evaluator;
#   kind "port":      oracle/add_oracle.py (the cited restatement) when nothing was staged.
# --------------------------------------------------------------------------------------------------
SEARCHED_DENSE_C2 = ([1, 2, 2, 2, 3, 2, 2, 1, 1, 1, 1, 2], [5], 0)      # eval.py:42-57 (network_arch, C_index, low_level_layer)


def cpu_synthetic_batch(n, h, w, seed=1234):
    """The SURVEY §8d synthetic batch (same generator calls as add_b200.synthetic_batch), built with torch only."""
    import torch
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, h, w, generator=g)
    gt = torch.randint(0, 19, (n, h, w), generator=g, dtype=torch.int64)
    gt[torch.rand(n, h, w, generator=g) < 0.1] = 255
    return x, gt


class CpuReference:
    """One image at a time through dynamic_inference + argmax + confusion matrix, like eval.py:195-221."""

    def __init__(self):
        import torch
        torch.set_num_threads(os.cpu_count() or 1)
        self.torch = torch
        ref = ROOT / "baseline" / "_ref"
        na, ci, low = SEARCHED_DENSE_C2
        if (ref / "modeling" / "ADD.py").exists() and (ref / "utils" / "metrics.py").exists():
            import numpy as np
            from types import SimpleNamespace
            sys.path.insert(0, str(ref))
            from modeling.ADD import ADD as RefADD, EDM as RefEDM       # noqa: E402  (the staged, unmodified reference)
            from utils.metrics import Evaluator as RefEvaluator          # noqa: E402
            # ADD.py:380,436 call torch.cuda.synchronize() unconditionally; this arm runs on CPU tensors, so the call is
            # made a no-op (it would fail without a GPU and create an idle CUDA context with one)
            torch.cuda.synchronize = lambda *a, **k: None
            cell = np.load(ref / "searched_arch" / "autodeeplab" / "genotype.npy")
            torch.manual_seed(1)                                         # eval.py:276,302
            self.model = RefADD(na, ci, cell, 19, SimpleNamespace(F=20, B=5, sync_bn=False), low).eval()
            torch.manual_seed(203)
            self.edm = RefEDM().eval()
            self.evaluator = RefEvaluator(19)
            self.kind = "reference"
            self.what = "unmodified reference (baseline/_ref): modeling.ADD.ADD.dynamic_inference + argmax + utils.metrics.Evaluator.add_batch"
        else:
            from oracle import add_oracle as orc
            self.orc = orc
            self.arch = orc.Arch(na, ci, low_level_layer=low)
            self.sd = orc.init_state_dict(self.arch, 1)
            self.edm_sd = orc.init_edm_state_dict(203)
            self.kind = "port"
            self.what = "oracle/add_oracle.py port of ADD.dynamic_inference + argmax + confusion matrix"

    def image(self, x1, gt1, threshold):
        torch = self.torch
        with torch.no_grad():
            if self.kind == "reference":
                y, ee, _, _ = self.model.dynamic_inference(x1, threshold, 'edm', self.edm)
                self.evaluator.add_batch(gt1, torch.argmax(y, 1))
                return ee
            y, ee, _ = self.orc.add_dynamic_inference(self.sd, self.arch, x1, threshold, 'edm', self.edm_sd)
            self.orc.generate_matrix(gt1.numpy(), torch.argmax(y, 1).numpy())
            return ee

    def warm(self):
        xs, gs = cpu_synthetic_batch(1, 128, 256, seed=5)
        self.image(xs, gs, 1e30)                                         # thread-pool / allocator warm-up (small)


def run_cpu_sample(a, n_images: int, exits: list) -> dict:
    """Time `n_images` images, image j forced to exit early iff exits[j] (same 50 % mix as the GPU arm)."""
    ref = CpuReference()
    x, gt = cpu_synthetic_batch(max(min(n_images, 2), 1), a.height, a.width)
    ref.warm()
    t0 = time.perf_counter()
    for j in range(n_images):
        thr = 1e30 if exits[j % len(exits)] else -1e30
        ref.image(x[j % x.shape[0]:j % x.shape[0] + 1], gt[j % x.shape[0]:j % x.shape[0] + 1], thr)
    dt = time.perf_counter() - t0
    return {"value": n_images / dt, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": ref.kind,
            "sample": f"{n_images} image(s) {a.height}x{a.width} fp32, {ref.what}, "
                      f"early-exit pattern {[int(exits[j % len(exits)]) for j in range(n_images)]}, {dt:.1f} s",
            "seconds": dt}


def main_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ref = CpuReference()
    x, gt = cpu_synthetic_batch(2, a.height, a.width)
    ref.warm()
    # one step = the b200 arm's batch: a.batch images, one dynamic_inference call each (the reference's gate is batch-1,
    # ADD.py:421), half of them forced to exit early (the b200 arm's median threshold gives the same mix).  The whole
    # `--steps K --warmup W` run has to end within a few minutes on whatever host cores the box has: one early-exit +
    # one full-depth image are timed first, and when (K + W) full batches would not fit --ref-budget-s the step becomes
    # a BOUNDED SAMPLE of the batch (an even number of images, same 50 % mix; images/s is unaffected by the sample size)
    t0 = time.perf_counter()
    ref.image(x[0:1], gt[0:1], 1e30)
    ref.image(x[1:2], gt[1:2], -1e30)
    t_pair = time.perf_counter() - t0
    n_step = a.batch
    fit = int(a.ref_budget_s / max((a.warmup + a.steps) * t_pair / 2, 1e-9))
    if fit < n_step:
        n_step = max(2, fit - fit % 2)
    times = []
    for s in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        for j in range(n_step):
            ref.image(x[j % 2:j % 2 + 1], gt[j % 2:j % 2 + 1], 1e30 if j % 2 == 0 else -1e30)
        if s >= a.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    val = n_step * len(times) / total
    cores = os.cpu_count() or 1
    sample = (f"{n_step} images {a.height}x{a.width} per step ("
              + ("the b200 arm's batch" if n_step == a.batch else
                 f"a bounded sample of the b200 arm's batch of {a.batch}: {a.warmup + a.steps} full batches would exceed the {a.ref_budget_s:.0f} s budget at {t_pair / 2:.2f} s per image")
              + f"), one batch-1 dynamic_inference call per image, alternating early-exit / full-depth, {ref.what}, fp32, {cores} host threads")
    line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_name(a),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": ref.kind, "sample": sample},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "images_per_step": n_step, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def cpu_reference_setup(a):
    """torch_cuda comparator only: oracle + weights built without the product."""
    import torch
    from oracle import add_oracle as orc
    na, ci, low = SEARCHED_DENSE_C2
    arch = orc.Arch(na, ci, low_level_layer=low)
    return orc, orc.init_state_dict(arch, 1), orc.init_edm_state_dict(203), arch


def main_torch_cuda(a):
    """Informational comparator ("also report stock PyTorch-CUDA images/sec as the GPU reference to beat"):
    the oracle restatement of ADD.dynamic_inference — plain torch.nn.functional calls, i.e. the cuDNN / ATen kernels
    the reference itself would run on this GPU — per image like eval.py:195-221 (batch 1, host gate decision, argmax,
    bincount confusion matrix), alternating early-exit / full-depth images.  Two passes: fp32 (TF32 off) and bf16
    weights + activations, both channels_last.  None of libadd_b200 is on this path."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    orc, sd, edm_sd, arch = cpu_reference_setup(a)
    dev = torch.device("cuda:0")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    x, gt = cpu_synthetic_batch(2, a.height, a.width)
    res = {}
    for name, dt in (("f32", torch.float32), ("bf16", torch.bfloat16)):
        def cast(v):
            v = v.to(dev)
            if v.is_floating_point():
                v = v.to(dt)
                if v.dim() == 4:
                    v = v.contiguous(memory_format=torch.channels_last)
            return v
        sdd = {k: cast(v) for k, v in sd.items()}
        edd = {k: cast(v) for k, v in edm_sd.items()}
        xd = [cast(x[j:j + 1]) for j in range(2)]
        gd = [gt[j:j + 1].to(dev) for j in range(2)]

        def one(j, thr):
            with torch.no_grad():
                y, ee, conf = orc.add_dynamic_inference(sdd, arch, xd[j], thr, 'edm', edd)
                pred = torch.argmax(y, 1)
                m = (gd[j] >= 0) & (gd[j] < 19)
                return torch.bincount(19 * gd[j][m] + pred[m], minlength=361).view(19, 19)
        for s_ in range(max(a.warmup, 3) * 2):
            one(s_ % 2, 1e30 if s_ % 2 == 0 else -1e30)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        n_img = 2 * max(a.steps, 2)
        for s_ in range(n_img):
            one(s_ % 2, 1e30 if s_ % 2 == 0 else -1e30)
        torch.cuda.synchronize()
        res[name] = n_img / (time.perf_counter() - t0)
    # the same network without gating, batched (ADD.forward on a.batch images -> argmax -> bincount per exit): what stock
    # PyTorch achieves when it is allowed a full batch; compare with `bench.py --forward-only` of this repo
    fwd = {}
    xb, gb = cpu_synthetic_batch(a.batch, a.height, a.width)
    for name, dt in (("bf16", torch.bfloat16),):
        sdd = {k: (v.to(dev).to(dt).contiguous(memory_format=torch.channels_last) if v.dim() == 4 else
                   (v.to(dev).to(dt) if v.is_floating_point() else v.to(dev))) for k, v in sd.items()}
        xd = xb.to(dev).to(dt).contiguous(memory_format=torch.channels_last)
        gd = gb.to(dev)
        m = (gd >= 0) & (gd < 19)

        def fwd_step():
            with torch.no_grad():
                outs = orc.add_forward(sdd, arch, xd)
                return [torch.bincount(19 * gd[m] + torch.argmax(o, 1)[m], minlength=361) for o in outs]
        for _ in range(3):
            fwd_step()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(max(a.steps, 2)):
            fwd_step()
        torch.cuda.synchronize()
        fwd[name] = a.batch * max(a.steps, 2) / (time.perf_counter() - t0)
    line = {"metric": METRIC, "value": res["bf16"], "unit": UNIT, "impl": "torch_cuda", "n_gpus": 1, "steps": a.steps,
            "forward_all_exits_batched_images_per_s": fwd,
            "warmup": max(a.warmup, 3), "higher_is_better": True, "dtype": "bf16", "data": "synthetic",
            "config": dict(workload_name(a), note="stock PyTorch (cuDNN/ATen) eager on cuda:0, batch 1 per call as in eval.py:195-221, "
                           "channels_last, alternating early-exit / full-depth images"),
            "images_per_s": res, "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(), "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------
def run_candidates(timeout_s: float = 150.0) -> dict:
    """Evaluator histogram kernels timed alone at the bench shape (8 x 1024 x 2048 int64 label / prediction pairs, 268 MB,
    three rotating sets = larger than L2): the default (per-warp, match.any) and the opt-in thread-private variant, in a
    child process.  Returns what the child measured, or why it could not."""
    import subprocess
    import tempfile
    out = Path(tempfile.gettempdir()) / f"add_b200_candidates_{os.getpid()}.json"
    note = ("opt-in kernels (add_confusion_set_impl(1)); not used by any default path, by `value` or by `e2e`; measured in a "
            "child process after the timed region")
    try:
        r = subprocess.run([sys.executable, str(ROOT / "tools" / "head_bench.py"), "20", "--variants", "--json", str(out)],
                           stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout_s, cwd=str(ROOT))
        res = json.loads(out.read_text()) if out.exists() else {}
        res.update({"rc": r.returncode, "note": note})
        if r.returncode != 0:
            res["stderr_tail"] = r.stderr[-600:]
        return res
    except subprocess.TimeoutExpired:
        res = json.loads(out.read_text()) if out.exists() else {}
        res.update({"rc": None, "error": f"child exceeded {timeout_s:.0f} s", "note": note})
        return res
    except Exception as e:          # the candidates must never take the bench line down
        return {"rc": None, "error": repr(e)[:300], "note": note}
    finally:
        try:
            out.unlink()
        except OSError:
            pass


def main_b200(a):
    # stdout carries exactly ONE JSON line: anything libraries print there meanwhile (NCCL's "NCCL version ..." banner
    # under torchrun) is diverted to stderr by pointing fd 1 at fd 2 until the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — add_b200 has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import __graft_entry__ as g
    if rank == 0:
        g.build()
    if world > 1:
        dist.barrier()
    import add_b200
    from add_b200 import runtime as rt

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    net = add_b200.build_add("searched-dense", 2, 20, seed=1).to(dev)
    net.set_precision(a.precision)
    net.use_cuda_graph = not a.no_graph
    torch.manual_seed(203)
    edm = add_b200.EDM().eval().to(dev)
    B, H, W = a.batch, a.height, a.width
    if a.image_dtype == "uint8":
        # the batch as the PNG decoder delivers it (uint8 HWC); x_host = the reference loader's normalisation of it
        img_host, x_host, gt_host = add_b200.synthetic_batch_u8(B, H, W, seed=1234 + rank, pin=True)
    else:
        x_host, gt_host = add_b200.synthetic_batch(B, H, W, seed=1234 + rank, pin=True)
        img_host = x_host
    x_dev, gt_dev = x_host.to(dev), gt_host.to(dev)

    # gate values for this batch → threshold between the two middle values (half the images exit early)
    _, _, confs = net.dynamic_evaluate(x_dev, gt_dev, -1e30, edm, a.exit_mode)
    vals = sorted(float(c) for c in confs)
    thr = 0.5 * (vals[B // 2 - 1] + vals[B // 2]) if B > 1 else vals[0] + 1.0
    if world > 1 and not a.per_rank_threshold:
        # ONE threshold for the whole job (rank 0's batch median), fixed before the timed region: the ranks' batches
        # differ (seed 1234 + rank), so their early-exit counts differ and the job runs at the pace of the slowest rank
        t = torch.tensor([thr], device=dev, dtype=torch.float64)
        dist.broadcast(t, 0)
        thr = float(t.item())
    cm0, flags0, _ = net.dynamic_evaluate(x_dev, gt_dev, thr, edm, a.exit_mode)
    cm0 = cm0.clone()
    launches_per_step = net.last_dynamic_launches + 0

    def step_resident():
        return net.dynamic_evaluate(x_dev, gt_dev, thr, edm, a.exit_mode)

    # resident arm: the batch lives in HBM in three buffers that are cycled through add_b200.ResidentPipeline — the
    # trunk of step i+1 is enqueued before the host reads step i's gate values, so the gate's host round trip
    # (~0.15 ms per step, tools/gate_bubble.py) does not leave the GPU idle.  --no-pipeline: one blocking call per step.
    res_bufs = [(x_dev, gt_dev)] if a.no_pipeline else [(x_dev.clone(), gt_dev.clone()) for _ in range(3)]
    rpipe = add_b200.ResidentPipeline(net, edm, thr, a.exit_mode)

    def run_resident(steps):
        out = None
        if a.no_pipeline:
            for _ in range(steps):
                out = step_resident()[0]
            return out
        for out, _ in rpipe.evaluate(res_bufs[i % len(res_bufs)] for i in range(steps)):
            pass
        return out

    # e2e: the public host-fed loop (add_b200.HostPipeline = eval.py's `for batch in loader` loop): every step's
    # images + labels come from pinned HOST memory (H2D inside the timed region, double-buffered on a copy
    # stream so batch i+1's copy overlaps batch i's compute) and its confusion matrices go back to the host.
    # Two feeds: (1) the batch as the PNG decoder delivers it — uint8 HWC images + uint8 labels, normalised on the
    # device (the default `e2e`); (2) the tensors the reference's loader hands to eval.py:175 — normalised fp32 NCHW
    # images + int64 labels, 5x the bytes (`e2e_fp32_feed`).
    def make_e2e(image_dtype, label_dtype):
        pipe_ = add_b200.HostPipeline(net, edm, thr, a.exit_mode)
        img_ = img_host if image_dtype == "uint8" else x_host
        gt_ = gt_host.to(torch.uint8).pin_memory() if label_dtype == "uint8" else gt_host   # 255 (ignore) fits uint8

        def run(steps):
            out = None
            for out, _ in pipe_.evaluate((img_, gt_) for _ in range(steps)):
                pass
            return out
        return pipe_, run

    def time_e2e(image_dtype, label_dtype):
        pipe_, run = make_e2e(image_dtype, label_dtype)
        run(pipe_.depth + 1)                                     # warm-up: every slot's plans recorded and captured, pinned result buffers
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        h2d0, d2h0 = pipe_.h2d_bytes, pipe_.d2h_bytes
        e0.record()
        cm_host = run(a.steps)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t_ = torch.tensor([ms], device=dev)
            dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            ms = float(t_.item())
        assert torch.equal(cm_host[0], cm0.cpu()), "e2e result differs from the resident-input result"
        return dict(value=world * B * a.steps / (ms / 1e3), unit=UNIT, h2d_bytes_per_step=(pipe_.h2d_bytes - h2d0) // a.steps,
                    d2h_bytes_per_step=(pipe_.d2h_bytes - d2h0) // a.steps + B * 4, ms_per_step=ms / a.steps,
                    host_label_dtype=label_dtype, host_image_dtype=image_dtype)

    if a.forward_only:
        # informational: ADD.evaluate (every exit, no gating: forward -> argmax -> confusion matrix per exit) on the
        # resident batch — the counterpart of `--impl torch_cuda`'s batched forward number
        for _ in range(3):
            net.evaluate(x_dev, gt_dev)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.steps):
            net.evaluate(x_dev, gt_dev)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps({"metric": "ADD 1024x2048 forward (all exits) images/sec", "value": B * a.steps / (ms / 1e3), "unit": UNIT,
                          "impl": "b200", "mode": "forward_only", "ms_per_step": ms / a.steps, "dtype": a.precision,
                          "config": workload_name(a)}), flush=True)
        return 0

    if a.profile_step:
        for _ in range(max(a.warmup, 3)):
            step_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step_resident()
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return 0

    def timed(run, steps, warmup, sample_clocks=False):
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.start()
        run(warmup)
        barrier()
        if sampler:
            sampler.mark()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = run(steps)
        e1.record()
        barrier()
        if sampler:
            sampler.mark()
        clocks = sampler.stop() if sampler else None
        ms = ms_local = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        assert torch.equal(last, cm0), "pipelined resident result differs from the single-call result"
        return ms, ms_local, clocks

    ms_res, ms_res_local, clocks = timed(run_resident, a.steps, max(a.warmup, 3), sample_clocks=True)
    if a.image_dtype == "fp32":
        e2e_main = time_e2e("fp32", a.label_dtype)
        e2e_fp32 = None
    else:
        e2e_main = time_e2e("uint8", a.label_dtype)
        e2e_fp32 = None if a.no_fp32_feed else time_e2e("fp32", "int64")
    # per-rank view of the resident arm: this rank's own step time and early-exit count (imbalance under one threshold)
    rank_ms, rank_exits = [ms_res_local / a.steps], [int(sum(flags0))]
    if world > 1:
        g_ms = [None] * world
        dist.all_gather_object(g_ms, (ms_res_local / a.steps, int(sum(flags0))))
        rank_ms, rank_exits = [v[0] for v in g_ms], [v[1] for v in g_ms]
    value = world * B * a.steps / (ms_res / 1e3)

    line = None
    if rank == 0:
        # ---- roofline: per-launch CUDA-event times over one full step (the plans the timed step replays), grouped into
        # kernel INSTANCES — launches of one kernel with the same tag and the same algorithmic work (= same shapes).
        # One instance per bound class is reported: the dominant tensor-bound instance (arithmetic intensity above the
        # bf16 ridge, FLOPs / time against the measured sustained bf16 peak) and the dominant HBM-bound instance
        # (algorithmic bytes / time against the measured copy bandwidth); the top-level keys are those of whichever
        # of the two takes more of the step.  `whole_step`: total algorithmic FLOPs / ms_per_step against the peak.
        runner = next(v for k, v in net._plans.items() if k[0] == "edm" and k[5] == "evaluate")
        rows = []
        for plan in runner.last_plans:                 # exactly the plans the timed step replays
            rows += plan.profile()
        peaks = {}
        pk = ROOT / "MEASURED_PEAKS.json"
        if pk.exists():
            peaks = json.loads(pk.read_text())
        # per-instance fractions: every launch is timed by its own CUDA-event pair in a ~10 ms eager replay at full
        # clocks (no power capping builds up) = "a kernel timed alone" -> the BURST bf16 figure; the whole-step
        # fraction (back-to-back timed steps) is quoted against the SUSTAINED one
        hbm_peak, tc_peak = peaks.get("hbm_gbs", 6650.0), peaks.get("bf16_tflops_sustained", 1400.0)
        tc_burst = peaks.get("bf16_tflops", tc_peak)
        src = ("MEASURED_PEAKS.json (hbm_gbs; bf16_tflops = burst for the per-launch-timed instances, bf16_tflops_sustained for whole_step)"
               if pk.exists() else "fallback (B200_PROFILING.md)")
        ridge = tc_peak * 1e12 / (hbm_peak * 1e9)
        agg, inst = {}, {}
        for r in rows:
            d = agg.setdefault(r["kernel"], dict(ms=0.0, flops=0, bytes=0, launches=0))
            d["ms"] += r["ms"]; d["flops"] += r["flops"]; d["bytes"] += r["bytes"]; d["launches"] += 1
            key = f'{r["tag"]}|{r["flops"]}|{r["bytes"]}'
            d = inst.setdefault(key, dict(kernel=r["kernel"], tag=r["tag"], ms=0.0, flops=r["flops"], bytes=r["bytes"], launches=0))
            d["ms"] += r["ms"]; d["launches"] += 1
        total_ms = sum(d["ms"] for d in agg.values())
        traffic_tab = {}
        tj = ROOT / "profiles" / "traffic.json"
        if tj.exists():                       # measured DRAM bytes per launch per instance (ncu launch list, committed)
            traffic_tab = json.loads(tj.read_text()).get("instances", {})

        def entry(key, d, bound):
            per_ms = d["ms"] / d["launches"]
            if bound == "tensor":
                ach, peak, unit, alg = d["flops"] / (per_ms / 1e3) / 1e12, tc_burst, "TFLOP/s", d["flops"]
            else:
                ach, peak, unit, alg = d["bytes"] / (per_ms / 1e3) / 1e9, hbm_peak, "GB/s", d["bytes"]
            tr = traffic_tab.get(key)
            return {"bound": bound, "achieved": ach, "peak": peak, "unit": unit, "frac": ach / peak,
                    **({"frac_of_sustained_peak": ach / tc_peak} if bound == "tensor" else {}),
                    "traffic": tr["traffic_per_launch"] if tr else None,
                    "traffic_source": "profiles/traffic.json (ncu dram__bytes_read+write per launch of this instance)" if tr else None,
                    "kernel": d["kernel"], "instance": d["tag"], "algorithmic_per_launch": alg,
                    "algorithmic_flops_per_launch": d["flops"], "algorithmic_bytes_per_launch": d["bytes"],
                    "launches": d["launches"], "avg_launch_ms": per_ms, "share_of_step": d["ms"] / total_ms}
        tens = {k: d for k, d in inst.items() if d["bytes"] > 0 and d["flops"] / d["bytes"] > ridge}
        mems = {k: d for k, d in inst.items() if d["bytes"] > 0 and d["flops"] / d["bytes"] <= ridge}
        classes = {}
        if tens:
            k = max(tens, key=lambda k: tens[k]["ms"]); classes["tensor"] = entry(k, tens[k], "tensor")
        if mems:
            k = max(mems, key=lambda k: mems[k]["ms"]); classes["hbm"] = entry(k, mems[k], "hbm")
        roof = dict(max(classes.values(), key=lambda c: c["share_of_step"]))
        # the SepConv halves are the largest kernel GROUP of the step (30 % of device time over 168 launches of ~20
        # shapes) though no single shape leads its class: their dominant instance is reported as well
        seps = {k: d for k, d in mems.items() if d["kernel"] == "sepconv_half_tc"}
        if seps:
            k = max(seps, key=lambda k: seps[k]["ms"]); classes["hbm_sepconv"] = entry(k, seps[k], "hbm")
            classes["hbm_sepconv"]["group_share_of_step"] = sum(d["ms"] for d in seps.values()) / total_ms
            classes["hbm_sepconv"]["group_launches"] = sum(d["launches"] for d in seps.values())
        step_flops = sum(r["flops"] for r in rows)
        step_bytes = sum(r["bytes"] for r in rows)
        roof.update({"peak_source": src, "classes": classes,
                     "whole_step": {"algorithmic_flops": step_flops, "algorithmic_bytes": step_bytes,
                                    "tflops": step_flops / (ms_res / a.steps / 1e3) / 1e12,
                                    "frac_of_bf16_sustained": step_flops / (ms_res / a.steps / 1e3) / 1e12 / tc_peak,
                                    "gbs": step_bytes / (ms_res / a.steps / 1e3) / 1e9,
                                    "frac_of_hbm": step_bytes / (ms_res / a.steps / 1e3) / 1e9 / hbm_peak,
                                    "note": "every launch of the step: op-granularity algorithmic work (SURVEY 8d) / measured ms_per_step"},
                     "families": {k: {"ms": round(v["ms"], 3), "launches": v["launches"],
                                      "tflops": round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 2),
                                      "gbs": round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 1)} for k, v in
                                  sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}})
        if a.profile_out:
            Path(a.profile_out).write_text(json.dumps(rows, indent=0))
        cpu = None
        if world == 1 and not a.no_cpu_baseline:
            cpu = run_cpu_sample(a, a.cpu_images, [True, False])   # same 50 % early-exit mix as the GPU arm
        # ---- candidates: kernels written after the round's GPU budget was spent and therefore NOT on any default path.
        # They are measured here — after every number of this line has been taken — in a SEPARATE process (its own CUDA
        # context, a hard timeout: nothing it does can touch the main measurement), next to the default kernel they
        # would replace, with a bit-identity check between the two.  tools/head_bench.py --variants.
        candidates = None
        if world == 1 and not a.no_candidates:
            candidates = run_candidates()
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
                "ms_per_step": ms_res / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": a.precision, "data": "synthetic", "config": dict(workload_name(a), parallelism=f"batch-shard dp{world}",
                                                                           early_exit_flags=flags0, edm_threshold=thr,
                                                                           cuda_graph=not a.no_graph,
                                                                           resident_pipeline=(None if a.no_pipeline else
                                                                                              "3 resident input buffers cycled; trunk of step i+1 enqueued before the host reads step i's gate values"),
                                                                           tensor_core_path=bool(rt.tc_available())),
                "e2e": {**e2e_main,
                        "api": "add_b200.HostPipeline.evaluate (pinned host batches over 3 slots: H2D of batch i+2 and the trunk of batch i+1 in flight while the host decides batch i's exits)"},
                "e2e_fp32_feed": e2e_fp32,
                "ranks": {"threshold": "per-rank batch median" if (a.per_rank_threshold or world == 1) else "global (rank 0's batch median, broadcast)",
                          "ms_per_step": rank_ms, "early_exits_of_batch": rank_exits,
                          "slowest_rank_penalty": max(rank_ms) / (sum(rank_ms) / len(rank_ms))},
                "gpu_launches": launches_per_step * a.steps, "gpu_launches_per_step": launches_per_step,
                "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "candidates": candidates}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    args = parse()
    sys.exit(main_reference(args) if args.impl == "reference" else main_torch_cuda(args) if args.impl == "torch_cuda"
             else main_b200(args))
