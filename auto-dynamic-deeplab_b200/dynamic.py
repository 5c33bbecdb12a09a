"""Early-exit inference (reference ADD.py:379-488) as segmented launch plans.

The reference decides per image on the host (`if confidence_value > threshold`, a device sync) and
is batch-1 only.  Here the network is recorded as *segments* — trunk up to each exit (+ the EDM
gate), one early-exit head per exit, the final trunk + head — and the host replays only what the
gate selects.  For the EDM gate a whole batch is gated per image: after each gate the exiting
images are gathered into an early-exit head plan, the continuing images are compacted (whole-image
slab gather, `add_gather_images`) into the next segment's plan, so exited images stop consuming
later layers.  Plans are keyed by (segment, image count) and built lazily.

Quirks reproduced (SURVEY Q3–Q5): the early exit uses the 2^-L `aspp_size` (the feature is
bilinearly up-sampled ×4 before ASPP), EDM's in-place ReLU is visible to the exit's resize and to
later cells, the last exit never resizes, EDM exits when value <= threshold.  `exit_mode='forward'`
is a labelled DEVIATION that sizes the early exit like `ADD.forward` does (2^-(L+2)).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from . import runtime as rt
from .runtime import Builder, Plan, View, RELU_IN
from .operations import _confidence


# ------------------------------------------------------------------------------------------------
# trunk state transfer between plans
# ------------------------------------------------------------------------------------------------

def _state_items(st: dict) -> List[Tuple[str, View]]:
    """Deterministic list of the trunk-state tensors later cells / heads read (ADD.py:388-412)."""
    items = [(f"two{i}", v) for i, v in enumerate(st["two"]) if v is not None]
    items += [(f"dense{i}", v) for i, v in enumerate(st["dense"])]
    if st.get("cur") is not None:
        items.append(("cur", st["cur"]))
    if st.get("low_cat") is not None:
        items.append(("low_cat", st["low_cat"]))
    return items


def _alloc_state_like(b: Builder, st: dict, m: int, need_two: bool) -> dict:
    """Fresh buffers for `m` images mirroring state `st`; aliasing between entries is preserved.
    `two_last_inputs` is only carried while cells < 3 remain (ADD.py:395-400)."""
    made: Dict[int, View] = {}

    def like(v: View) -> View:
        key = id(v.buf)
        if key not in made:
            made[key] = View(b.raw((m,) + tuple(v.buf.shape[1:]), v.buf.dtype))
        return View(made[key].buf, v.c_off, v.c, v.relud)

    new = dict(two=[like(v) if need_two else None for v in st["two"]], dense=[like(v) for v in st["dense"]],
               cur=like(st["cur"]) if st.get("cur") is not None else None,
               low_cat=like(st["low_cat"]) if st.get("low_cat") is not None else None, size=st["size"])
    return new


def _low_slot(cat: View) -> View:
    """The low-level slice of the decoder's concat buffer (decoder.py:26: cat(x[256], low_level[48]))."""
    from .decoder import ASPP_C, LOW_LEVEL_C
    return cat.slice(ASPP_C, LOW_LEVEL_C)


def _emit_state_gather(b: Builder, src: dict, dst: dict, idx: torch.Tensor) -> None:
    done = set()
    for (name, s), (_, d) in zip(_state_items(src), _state_items(dst)):
        if id(d.buf) in done:
            continue
        done.add(id(d.buf))
        if name == "low_cat":
            # only the low-level slot of the decoder's concat buffer is live here (the ASPP part is written later)
            b.gather_view(_low_slot(s), _low_slot(d), idx, "dynamic.gather.low_cat")
            continue
        b.gather_images(s.buf, d.buf, idx, "dynamic.gather." + name)


class _Segment:
    """Trunk cells (after exit k-1) .. exit k, plus the EDM gate when exit k is not the last layer.
    For the last segment the final head is part of the plan."""

    def __init__(self, runner: "_EdmRunner", k: int, m: int, prev: Optional["_Segment"]):
        net = runner.net
        self.k, self.m = k, m
        b = Builder(runner.device, runner.dtype, record=True)
        self.builder = b
        H, W = runner.H, runner.W
        first = 0 if k == 0 else runner.exits[k - 1] + 1
        is_last = k == len(runner.exits)
        last = net.num_net - 1 if is_last else runner.exits[k]
        self.idx = b.raw((m,), torch.int32, zero=True)   # valid ids even before the host fills them
        self.gather: Optional[Plan] = None
        if k == 0:
            if runner.bound is not None:
                self.x_static = runner.bound[0]
                b.keep.append(self.x_static)
            else:
                self.x_static = b.raw((m, 3, H, W), torch.float32)
            st: dict = {}
        else:
            st = _alloc_state_like(b, prev.state, m, need_two=first <= 2)
            start = len(b.launches)
            src = dict(prev.state)
            if first > 2:
                src["two"] = [None, None]
            _emit_state_gather(b, src, st, self.idx)
            self.gather_range = (start, len(b.launches))
        start = len(b.launches)
        net._emit_trunk(b, self.x_static if k == 0 else None, first, last, st)
        y = net._feature(st, last)
        self.conf = None
        self.out = None
        if not is_last:
            self.conf = runner.edm.emit_edm(b, y)
            # EDM.forward's in-place ReLU (ADD.py:516-519) mutates the feature every later reader sees (Q4)
            yr = b.alloc(y.n, y.h, y.w, y.c)
            b.bilinear(y, yr, RELU_IN, "EDM.inplace_relu")
            if last > 2:
                st["cur"] = yr
            else:
                st["two"][1] = yr
                if last == 2:
                    st["cur"] = yr     # x aliases two_last_inputs[1] at i == 2 (ADD.py:399-400)
        else:
            # last exit: never resized in the EDM path (ADD.py:433-435)
            logits = net._emit_exit_lowres(b, y, st, last, runner.aspp_size, 0, resize=False)
            self.out = runner.emit_head_output(b, logits, m, self)
        self.state = st
        self.main = Plan(b, start, len(b.launches))
        if k > 0:
            self.gather = Plan(b, *self.gather_range)
        if net.use_cuda_graph:
            if k > 0 and not is_last or (k > 0 and _TRUNK_PCT != 100):
                _capture_with_pct(self.main, _TRUNK_PCT)
            else:
                self.main.capture()
            if self.gather is not None:
                self.gather.capture()

    @property
    def n_launches(self) -> int:
        return self.main.n_launches + (self.gather.n_launches if self.gather is not None else 0)


class _Head:
    """Early-exit head k for `m` gathered images: [resize →] [conv_aspp →] ASPP → decoder → output."""

    def __init__(self, runner: "_EdmRunner", k: int, m: int, seg: _Segment):
        net = runner.net
        b = Builder(runner.device, runner.dtype, record=True)
        self.builder = b
        self.idx = b.raw((m,), torch.int32, zero=True)   # valid ids even before the host fills them
        i = runner.exits[k]
        y_src = net._feature(seg.state, i)
        y = View(b.raw((m,) + tuple(y_src.buf.shape[1:]), y_src.buf.dtype))
        low_src = seg.state["low_cat"]
        low = View(b.raw((m,) + tuple(low_src.buf.shape[1:]), low_src.buf.dtype))
        b.gather_images(y_src.buf, y.buf, self.idx, "dynamic.gather.exit_feature")
        b.gather_view(_low_slot(low_src), _low_slot(low), self.idx, "dynamic.gather.low_cat")
        # conv_aspp_iter == k: every earlier exit was skipped (ADD.py:422)
        logits = net._emit_exit_lowres(b, y, dict(low_cat=low), i, runner.early_aspp_size, k, True, True)
        self.out = runner.emit_head_output(b, logits, m, self)
        self.main = Plan(b)
        if net.use_cuda_graph:
            _capture_with_pct(self.main, _HEAD_PCT)

    @property
    def n_launches(self) -> int:
        return self.main.n_launches


import os as _os
_OVERLAP_HEADS = _os.environ.get("ADD_OVERLAP_HEADS", "1") != "0"
# share of the persistent kernels' CTAs given to the early-exit head plan / to the trunk plan that runs beside it
_HEAD_PCT = int(_os.environ.get("ADD_HEAD_PCT", "100"))
_TRUNK_PCT = int(_os.environ.get("ADD_TRUNK_PCT", "100"))


def _capture_with_pct(plan, pct: int) -> None:
    from ._lib import lib, check
    check(lib.add_set_persistent_grid_pct(pct), "set_persistent_grid_pct")
    try:
        plan.capture()
    finally:
        check(lib.add_set_persistent_grid_pct(100), "set_persistent_grid_pct")


class _EdmRunner:
    """Batched, per-image EDM-gated inference for one (input shape, precision, output mode)."""

    def __init__(self, net, shape, device, precision: str, edm, mode: str, exit_mode: str,
                 bound: Optional[Tuple[torch.Tensor, Optional[torch.Tensor]]] = None,
                 label_dtype: torch.dtype = torch.int64):
        """bound = (x, target): record the plans directly on THESE device tensors (the caller promises to refill the
        same buffers for every call, as HostPipeline's slots do) — no copy into private static inputs."""
        self.generation = rt.generation()
        self.bound = bound
        self.net, self.edm, self.mode = net, edm, mode
        self.device, self.dtype = device, rt.act_dtype(precision)
        self.n, _, self.H, self.W = shape
        self.nc = net._num_classes
        self.aspp_size = net._aspp_size((self.H, self.W), net.network_arch[-1])          # ADD.py:383-384
        if exit_mode == "reference":
            self.early_aspp_size = self.aspp_size
        elif exit_mode == "forward":     # DEVIATION: size the early exit like ADD.forward (ADD.py:279-280)
            self.early_aspp_size = net._aspp_size((self.H, self.W), net.network_arch[-1] + 2)
        else:
            raise ValueError(exit_mode)
        self.exits = [i for i in range(net.num_net) if i in net.C_index and i != net.num_net - 1]
        self.segments: Dict[Tuple[int, int], _Segment] = {}
        self.heads: Dict[Tuple[int, int], _Head] = {}
        self.gt_full: Optional[torch.Tensor] = None
        self.cm_full: Optional[torch.Tensor] = None
        if mode == "evaluate":
            self.cm_store = torch.zeros((self.n + 1, self.nc, self.nc), dtype=torch.int64, device=device)   # row N: spare
            self.cm_full = self.cm_store[:self.n]
            # labels: the reference's int64 tensors, or uint8 (the PNG bytes: 8x less gather / histogram traffic)
            self.label_dtype = label_dtype
            self.gt_full = (bound[1] if bound is not None else
                            torch.empty((self.n, self.H, self.W), dtype=label_dtype, device=device))
        self.last_launches = 0
        self._side: Optional[torch.cuda.Stream] = None
        # pinned staging for the gate's host round trip (N floats down, a few index vectors up): no pageable copies
        self._conf_pin = torch.empty(self.n, dtype=torch.float32).pin_memory()
        self._idx_pin = torch.empty((16, 2 * self.n), dtype=torch.int32).pin_memory()
        self._idx_np = self._idx_pin.numpy()
        self._idx_slot = 0

    # ---- per-plan output tail -------------------------------------------------------------------
    def emit_head_output(self, b: Builder, logits: View, m: int, owner) -> torch.Tensor:
        if self.mode == "logits":
            out = b.raw((m, self.nc, self.H, self.W), torch.float32)
            b.upsample_logits(logits, out, self.H, self.W, "ADD.upsample_logits")
            return out
        gt = b.raw((m, self.H, self.W), self.label_dtype)
        # ORIGINAL image ids of this plan's rows, twice: [0, m) drives the label gather, [m, 2m) the scatter of the
        # confusion matrices into the [N, nc, nc] result of the original batch (cm_row_index of
        # add_upsample_argmax_fwd: no stacking of per-plan outputs afterwards).  Until the host uploads real ids (the
        # warm-up run of Plan.capture) the gather reads image 0 and the scatter goes to the spare row N of cm_store,
        # so recording a plan never disturbs results that other plans have already written for this batch.
        owner.idx_gt = torch.tensor([0] * m + [self.n] * m, dtype=torch.int32).to(self.device)    # host-built: a memcpy
        b.keep.append(owner.idx_gt)
        b.gather_images(self.gt_full, gt, owner.idx_gt[:m], "dynamic.gather.gt")
        b.upsample_argmax(logits, self.H, self.W, gt, None, self.cm_store, None, "ADD.upsample_argmax_cm",
                          cm_rows=owner.idx_gt[m:])
        return self.cm_full

    def segment(self, k: int, m: int, prev: Optional[_Segment]) -> _Segment:
        # a plan is recorded against the buffers of ONE specific source segment: key it by that segment's identity
        # (with >= 3 gated exits several segments share (k, m) and differ only in their lineage)
        key = (k, m, id(prev) if prev is not None else 0)
        if key not in self.segments:
            self.segments[key] = _Segment(self, k, m, prev)
        return self.segments[key]

    def head(self, k: int, m: int, seg: _Segment) -> _Head:
        key = (k, m, id(seg))
        if key not in self.heads:
            self.heads[key] = _Head(self, k, m, seg)
        return self.heads[key]

    def _put_idx(self, dst: torch.Tensor, values) -> None:
        """Asynchronous H2D of a short index vector through a rotating pinned row."""
        row = self._idx_slot
        self._idx_slot = (row + 1) % self._idx_pin.shape[0]
        m = len(values)
        self._idx_np[row, :m] = values
        dst.copy_(self._idx_pin[row, :m], non_blocking=True)

    # ---- one batch ------------------------------------------------------------------------------
    def run(self, x: torch.Tensor, threshold: float, target: Optional[torch.Tensor] = None):
        """Returns (outputs per image, exit flag per image, confidence per image).  outputs[j] is a
        [1,nc,H,W] fp32 logits tensor ('logits') or an int64 [nc,nc] confusion matrix ('evaluate');
        they alias plan buffers that the next call overwrites."""
        return self.finish(self.begin(x, threshold, target))

    # The gate is a host decision (the reference's implicit sync, ADD.py:421): after the trunk graph the host reads N
    # floats, picks the plans for the exiting / continuing images and launches them — ~0.15 ms per step during which
    # the GPU is idle (tools/gate_bubble.py).  begin() enqueues everything up to the first gate and returns; finish()
    # waits for that gate's values and enqueues the rest.  A caller that owns several runners (HostPipeline's slots)
    # calls begin() of batch i+1 BEFORE finish() of batch i, so the GPU works on the next trunk while the host decides.
    def begin(self, x: torch.Tensor, threshold: float, target: Optional[torch.Tensor] = None):
        gen = self._steps(x, threshold, target)
        try:
            return [gen, next(gen), None]                 # [coroutine, event of the pending gate, result]
        except StopIteration as stop:
            return [None, None, stop.value]

    def finish(self, state):
        gen, ev, result = state
        while gen is not None:
            ev.synchronize()                              # the gate values are in pinned host memory
            try:
                ev = next(gen)
            except StopIteration as stop:
                gen, result = None, stop.value
        return result

    def _steps(self, x: torch.Tensor, threshold: float, target: Optional[torch.Tensor] = None):
        """Coroutine behind begin()/finish(): yields a CUDA event after each gate's D2H copy is enqueued and expects
        to be resumed once that event has completed."""
        n = self.n
        outs: List[Optional[torch.Tensor]] = [None] * n
        flags = [0] * n
        confs: List[Optional[torch.Tensor]] = [None] * n
        launches = 0
        pending_side = False
        self.last_plans: List[Plan] = []
        if self.mode == "evaluate" and self.bound is None:
            self.gt_full.copy_(target, non_blocking=True)
        active = list(range(n))
        seg = self.segment(0, n, None)
        if self.bound is None:
            seg.x_static.copy_(x, non_blocking=True)
        for k in range(len(self.exits) + 1):
            if k > 0:
                seg.gather.run()
                self.last_plans.append(seg.gather)
            if k == len(self.exits) and self.mode == "evaluate":
                self._put_idx(seg.idx_gt, list(active) * 2)
            seg.main.run()
            self.last_plans.append(seg.main)
            launches += seg.n_launches
            if k == len(self.exits):
                for j, img in enumerate(active):
                    outs[img] = seg.out[j:j + 1] if self.mode == "logits" else self.cm_full[img]
                break
            # host decision = the reference's implicit sync (ADD.py:421): N floats come down through pinned memory
            m_act = len(active)
            self._conf_pin[:m_act].copy_(seg.conf, non_blocking=True)
            gate_ev = torch.cuda.Event()
            gate_ev.record(torch.cuda.current_stream(self.device))
            yield gate_ev
            vals = self._conf_pin[:m_act].tolist()
            ex = [j for j in range(m_act) if not (vals[j] > threshold)]
            co = [j for j in range(m_act) if vals[j] > threshold]
            conf = self._conf_pin[:m_act].clone()
            for j, img in enumerate(active):
                confs[img] = conf[j].view(1, 1)
            if ex:
                head = self.head(k, len(ex), seg)
                self._put_idx(head.idx, ex)
                if self.mode == "evaluate":
                    self._put_idx(head.idx_gt, [active[j] for j in ex] * 2)
                if co and _OVERLAP_HEADS:
                    # the early-exit head (throughput-bound: ASPP on the up-sampled map) and the remaining trunk
                    # (latency-bound small kernels) only READ this segment's state: replay them side by side
                    main = torch.cuda.current_stream(self.device)
                    if self._side is None:
                        self._side = torch.cuda.Stream(self.device)
                    self._side.wait_stream(main)
                    with torch.cuda.stream(self._side):
                        head.main.run()
                    pending_side = True
                else:
                    head.main.run()
                self.last_plans.append(head.main)
                launches += head.n_launches
                for jj, j in enumerate(ex):
                    outs[active[j]] = head.out[jj:jj + 1] if self.mode == "logits" else self.cm_full[active[j]]
                    flags[active[j]] = 1
            if not co:
                break
            nxt = self.segment(k + 1, len(co), seg)
            self._put_idx(nxt.idx, co)
            active = [active[j] for j in co]
            seg = nxt
        if pending_side:
            torch.cuda.current_stream(self.device).wait_stream(self._side)
        self.last_launches = launches
        return outs, flags, confs


def _get_runner(net, x: torch.Tensor, edm, mode: str, exit_mode: str, target: Optional[torch.Tensor] = None,
                bind_inputs: bool = False) -> _EdmRunner:
    prec = net.precision or rt.default_precision()
    label_dtype = torch.int64
    if target is not None:
        if target.dtype not in (torch.int64, torch.uint8):
            raise ValueError(f"labels must be int64 (metrics.py:34-39) or uint8, got {target.dtype}")
        label_dtype = target.dtype
    key = ("edm", tuple(x.shape), str(x.device), prec, id(edm), mode, exit_mode, bool(net.use_cuda_graph), label_dtype)
    bound = None
    if bind_inputs:
        if not (x.dtype == torch.float32 and x.is_contiguous() and (target is None or target.is_contiguous())):
            raise ValueError("bind_inputs needs contiguous fp32 NCHW images and contiguous int64 / uint8 labels")
        key = key + (x.data_ptr(), 0 if target is None else target.data_ptr())
        bound = (x, target)
    r = net._plans.get(key)
    if r is None or r.generation != rt.generation():
        r = _EdmRunner(net, tuple(x.shape), x.device, prec, edm, mode, exit_mode, bound, label_dtype)
        net._plans[key] = r
    return r


# ------------------------------------------------------------------------------------------------
# entropy / max gates: the exit head is evaluated first, then scored (ADD.py:440-488) — batch 1
# ------------------------------------------------------------------------------------------------

class _ScorePlan:
    def __init__(self, net, shape, device, precision: str):
        self.generation = rt.generation()
        n, _, H, W = shape
        assert n == 1
        nc = net._num_classes
        b = Builder(device, rt.act_dtype(precision), record=True)
        self.builder = b
        self.x_static = b.raw(shape, torch.float32)
        aspp_size = net._aspp_size((H, W), net.network_arch[-1])          # ADD.py:442-443
        st: dict = {}
        self.stages: List[Plan] = []
        self.outs: List[torch.Tensor] = []
        exits = [i for i in range(net.num_net) if i in net.C_index or i == net.num_net - 1]
        done = -1
        for k, i in enumerate(exits):
            start = len(b.launches)
            net._emit_trunk(b, self.x_static, done + 1, i, st)
            done = i
            y = net._feature(st, i)
            if not (y.h < aspp_size[0] or y.w < aspp_size[1]):
                raise NotImplementedError("non-EDM gate with a feature not smaller than aspp_size: the reference "
                                          "skips the head and scores the raw feature map (ADD.py:465-476)")
            logits = net._emit_exit_lowres(b, y, st, i, aspp_size, k, True, False)
            out = b.raw((1, nc, H, W), torch.float32)
            b.upsample_logits(logits, out, H, W, "ADD.upsample_logits")
            self.outs.append(out)
            self.stages.append(Plan(b, start, len(b.launches)))
        if net.use_cuda_graph:
            for p in self.stages:
                p.capture()


def _run_scored(net, x: torch.Tensor, threshold, confidence):
    prec = net.precision or rt.default_precision()
    ys, flags, confs, launches = [], [], [], 0
    for n in range(x.shape[0]):
        x1 = x[n:n + 1]
        key = ("scored", tuple(x1.shape), str(x1.device), prec, bool(net.use_cuda_graph))
        plan = net._plans.get(key)
        if plan is None or plan.generation != rt.generation():
            plan = _ScorePlan(net, tuple(x1.shape), x1.device, prec)
            net._plans[key] = plan
        plan.x_static.copy_(x1)
        conf_val, taken = None, len(plan.stages) - 1
        for k, stage in enumerate(plan.stages):
            stage.run()
            launches += stage.n_launches
            if k == len(plan.stages) - 1:
                break
            s = _confidence(plan.outs[k], threshold if confidence == 'max' else 2.0, net._num_classes)
            if confidence == 'entropy':
                conf_val = s[0].item()
                hit = conf_val < threshold
            else:
                conf_val = s[1].item()
                hit = conf_val > threshold
            if hit:
                taken = k
                break
        out = plan.outs[taken]
        ys.append(out.clone() if x.shape[0] > 1 else out)
        flags.append(0 if taken == len(plan.stages) - 1 else 1)
        confs.append(conf_val)
    net.last_dynamic_launches = launches
    return ys, flags, confs


# ------------------------------------------------------------------------------------------------
# entry points used by ADD
# ------------------------------------------------------------------------------------------------

def run_dynamic(net, x: torch.Tensor, threshold, confidence, edm, exit_mode: str = "reference"):
    if confidence not in ('edm', 'entropy', 'max'):
        raise ValueError(confidence)
    if confidence != 'edm':
        return _run_scored(net, x, threshold, confidence)
    if edm is False or edm is None:
        raise ValueError("confidence='edm' needs an EDM module (eval.py:110-112)")
    r = _get_runner(net, x, edm, "logits", exit_mode)
    outs, flags, confs = r.run(x, float(threshold))
    net.last_dynamic_launches = r.last_launches
    return outs, flags, confs


def begin_dynamic_evaluate(net, x: torch.Tensor, target: torch.Tensor, threshold, edm, exit_mode: str = "reference",
                           bind_inputs: bool = False):
    """First half of run_dynamic_evaluate: enqueue the trunk up to the first gate and return a handle."""
    r = _get_runner(net, x, edm, "evaluate", exit_mode, target, bind_inputs)
    return r, r.begin(x, float(threshold), target)


def finish_dynamic_evaluate(net, handle):
    """Second half: wait for the gate values, enqueue the exit heads / remaining trunk.  Returns (cm, flags, confs);
    cm [N,nc,nc] int64 ALIASES the runner's result buffer (every plan scatters its images' matrices into it): it is
    overwritten by the next batch begun on the same input buffers."""
    r, state = handle
    outs, flags, confs = r.finish(state)
    net.last_dynamic_launches = r.last_launches
    return r.cm_full, flags, confs


def run_dynamic_evaluate(net, x: torch.Tensor, target: torch.Tensor, threshold, edm, exit_mode: str = "reference",
                         bind_inputs: bool = False):
    """eval.py:195-221 for a batch, fused: per-image EDM gate → exit head → argmax → int64 confusion
    matrix, without materialising full-resolution logits.  Returns (cm int64 [N,nc,nc], flags, confs).
    bind_inputs: record the plans on x / target themselves (stable buffers refilled by the caller, e.g. the slots
    of HostPipeline) instead of copying into private static inputs."""
    r = _get_runner(net, x, edm, "evaluate", exit_mode, target, bind_inputs)
    outs, flags, confs = r.run(x, float(threshold), target)
    net.last_dynamic_launches = r.last_launches
    # the blocking call hands out a private copy (one 8*N*nc*nc-byte device memcpy); the begin / finish pair used by
    # the pipelines returns the runner's own result buffer, which the next batch on the same inputs overwrites
    return r.cm_full.clone(), flags, confs
