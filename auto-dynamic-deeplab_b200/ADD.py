"""The ADD network — B200 drop-in for `modeling/ADD.py` (Cell :14-116, ADD :118-500, EDM :502-525).

Same constructor, attribute names and 1998 state_dict keys as the reference.  `forward`,
`get_feature` and `dynamic_inference` do not execute an nn.Module graph: on first use for a given
input shape they *record* a static launch plan over preallocated NHWC buffers (concat = channel
slice writes, node sum = accumulate-into-slice, BN folded, ReLU on load/store) and afterwards
replay it — as a CUDA graph when `use_cuda_graph` is set.
"""
from __future__ import annotations

import ctypes
import time
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import runtime as rt
from .genotypes import PRIMITIVES
from .aspp_train import ASPP_train
from .decoder import Decoder, ASPP_C, LOW_LEVEL_C
from .operations import (AddModule, OPS, ReLUConvBN, FactorizedReduce, DoubleFactorizedReduce,
                         SynchronizedBatchNorm2d, _conv_holder, normalized_shannon_entropy, confidence_max)
from .runtime import Builder, ConvWeights, Plan, View, RELU_IN, RELU_OUT, ACCUMULATE, IN_RELUD
from ._lib import lib, check


def _scale_dimension(dim: int, scale: float) -> int:
    return int((float(dim) - 1.0) * scale + 1.0)


def executed_edges(cell_arch: np.ndarray, B: int):
    """Per step, the (state j, op index, primitive) edges Cell.forward really executes
    (reference ADD.py:97-110; ops are bound in ascending-branch order — SURVEY Q1)."""
    selected = set(int(v) for v in np.asarray(cell_arch)[:, 0])
    steps, offset, k, n_states = [], 0, 0, 2
    for _ in range(B):
        edges = []
        for j in range(n_states):
            if offset + j in selected:
                edges.append((j, k))
                k += 1
        steps.append(edges)
        offset += n_states
        n_states += 1
    return steps


class Cell(AddModule):
    """reference ADD.py:14-116."""

    def __init__(self, BatchNorm, B, prev_prev_C, prev_C, cell_arch, network_arch, C_out, downup_sample,
                 dense_in=False, dense_out=True):
        super().__init__()
        eps, momentum = 1e-5, 0.1
        self.cell_arch = cell_arch
        self.downup_sample = downup_sample
        self.B = B
        self.dense_in = dense_in
        self.dense_out = dense_out
        self.C_out = C_out
        if downup_sample == -1:
            self.preprocess = FactorizedReduce(prev_C, C_out, BatchNorm, eps=eps, momentum=momentum)
        else:
            self.preprocess = ReLUConvBN(prev_C, C_out, 1, 1, 0, BatchNorm, eps=eps, momentum=momentum, affine=True)
        if downup_sample == 1:
            self.scale = 2
        self._ops = nn.ModuleList()
        if dense_in:
            self.pre_preprocess = nn.ModuleList(
                ReLUConvBN(c, C_out, 1, 1, 0, BatchNorm, eps=eps, momentum=momentum, affine=True) for c in prev_prev_C)
            self.pre_preprocess_1x1 = ReLUConvBN(len(prev_prev_C) * C_out, C_out, 1, 1, 0, BatchNorm,
                                                 eps=eps, momentum=momentum, affine=True)
        else:
            self.pre_preprocess = ReLUConvBN(prev_prev_C, C_out, 1, 1, 0, BatchNorm, eps=eps, momentum=momentum, affine=True)
        if dense_out:
            self.dense_process = ReLUConvBN(C_out * B, C_out, 1, 1, 0, BatchNorm, eps=eps, momentum=momentum, affine=True)
        arch = cell_arch.numpy() if torch.is_tensor(cell_arch) else np.asarray(cell_arch)
        for row in arch:
            self._ops.append(OPS[PRIMITIVES[int(row[1])]](C_out, 1, BatchNorm, eps=eps, momentum=momentum, affine=True))
        self._steps = executed_edges(arch, B)
        # s0 / s1 may be stored post-ReLU only if every executed op starts with ReLU (sep/dil convs do, operations.py:33,47;
        # pools, skip_connect and none take the raw signed input)
        relu_first = {'sep_conv_3x3', 'sep_conv_5x5', 'dil_conv_3x3', 'dil_conv_5x5'}
        self._all_relu_first = all(PRIMITIVES[int(row[1])] in relu_first for row in arch)

    def _prepare(self):
        pass

    def scale_dimension(self, dim, scale):
        return _scale_dimension(dim, scale)

    def emit_cell(self, b: Builder, prev_prev, prev: View, concat_relud: bool = False) -> Tuple[View, Optional[View]]:
        """Emit the whole cell.  prev_prev: View (dense_in False) or list of Views.
        Returns (concat [N,h,w,B*C], dense [N,h,w,C] or None).
        concat_relud: the caller guarantees that every reader of the concat outside this cell starts with ReLU
        (next cell's ReLUConvBN / FactorizedReduce, low_level_conv — not a bilinear resize, not an exit); the node
        sums are then stored as relu(sum) by their last edge and all readers skip the ReLU-on-load pass."""
        C, n = self.C_out, prev.n
        temps: List[View] = []
        s1_in = prev
        if self.downup_sample == 1:  # ADD.py:71-77
            up = b.scratch(n, _scale_dimension(prev.h, 2), _scale_dimension(prev.w, 2), prev.c)
            b.bilinear(prev, up, 0, "Cell.up")
            temps.append(up)
            s1_in = up
        if self.downup_sample == -1:
            h, w = (s1_in.h - 1) // 2 + 1, (s1_in.w - 1) // 2 + 1
        else:
            h, w = s1_in.h, s1_in.w
        s1 = b.scratch(n, h, w, C)
        # s0 / s1 are read only by the cell's ops, which all start with ReLU (operations.py:33,47): store relu(s)
        # once (RELU_OUT) and let the five ops that read them skip their ReLU-on-load pass
        pre_flags = RELU_OUT if self._all_relu_first else 0
        concat_relud = concat_relud and self._all_relu_first
        self.preprocess.emit(b, s1_in, s1, pre_flags)
        s0 = b.scratch(n, h, w, C)
        temps += [s0, s1]
        if not self.dense_in:  # ADD.py:83-86
            src = prev_prev
            if src.h != h:
                r = b.scratch(n, h, w, src.c)
                b.bilinear(src, r, 0, "Cell.resize_pp")
                temps.append(r)
                src = r
            self.pre_preprocess.emit(b, src, s0, pre_flags)
        else:  # ADD.py:87-93
            k = len(prev_prev)
            cat = b.scratch(n, h, w, k * C)
            temps.append(cat)
            for i, src in enumerate(prev_prev):
                if src.h != h:
                    r = b.resized(src, h, w, "Cell.resize_dense")
                    self.pre_preprocess[i].emit(b, r, cat.slice(i * C, C), 0)
                else:
                    self.pre_preprocess[i].emit(b, src, cat.slice(i * C, C), 0)
            self.pre_preprocess_1x1.emit(b, cat, s0, pre_flags)
        concat = b.alloc(n, h, w, self.B * C)
        s0.relud = s1.relud = bool(pre_flags)
        states = [s0, s1] + [concat.slice(i * C, C) for i in range(self.B)]
        for i, edges in enumerate(self._steps):  # ADD.py:97-110
            dst = states[2 + i]
            if not edges:
                raise NotImplementedError("cell step without inputs (sum of empty list) is not supported")
            for e, (j, k) in enumerate(edges):
                last = e == len(edges) - 1
                self._ops[k].emit(b, states[j], dst, (ACCUMULATE if e > 0 else 0) | (RELU_OUT if (last and concat_relud) else 0))
            dst.relud = concat_relud              # later edges of this cell read relu(node sum) without a ReLU pass
        concat.relud = concat_relud
        dense = None
        if self.dense_out:
            dense = b.alloc(n, h, w, C)
            self.dense_process.emit(b, concat, dense, 0)
        for t in temps:
            b.release(t)
        return concat, dense

    def forward(self, prev_prev_input, prev_input):
        """Stand-alone call with the reference's signature/returns (ADD.py:69-116)."""
        self._check_eval()
        rt.require_cuda(prev_input)
        dtype = prev_input.dtype if prev_input.dtype in (torch.float32, torch.bfloat16) else torch.float32
        b = Builder(prev_input.device, dtype, record=False)
        pv = rt.as_nhwc_view(prev_input, b, dtype)
        if self.dense_in:
            ppv = [rt.as_nhwc_view(t, b, dtype) for t in prev_prev_input]
        else:
            ppv = rt.as_nhwc_view(prev_prev_input, b, dtype)
        concat, dense = self.emit_cell(b, ppv, pv)
        if self.dense_out:
            return prev_input, concat.nchw(), dense.nchw()
        return concat.nchw()


class _Stem(nn.Sequential):
    pass


class ADD(AddModule):
    """reference ADD.py:118-500."""

    def __init__(self, network_arch, C_index, cell_arch, num_classes, args, low_level_layer):
        super().__init__()
        BatchNorm = SynchronizedBatchNorm2d if args.sync_bn == True else nn.BatchNorm2d  # noqa: E712
        F_, B = args.F, args.B
        eps, momentum = 1e-5, 0.1
        self.args = args
        self.cell_arch = torch.from_numpy(np.asarray(cell_arch))
        self._num_classes = num_classes
        self.low_level_layer = low_level_layer
        self.network_arch = list(network_arch)
        self.num_net = len(network_arch)
        self.C_index = list(C_index)
        self.precision: Optional[str] = None     # None → runtime default ('fp32' | 'bf16')
        self.use_cuda_graph = False
        self._plans: Dict[tuple, "_NetPlan"] = {}

        self.decoder = Decoder(num_classes, BatchNorm)
        FB = F_ * B
        fm = {0: 1, 1: 2, 2: 4, 3: 8}
        self.stem0 = nn.Sequential(_conv_holder(3, 64, 3, 2, 1), BatchNorm(64, eps=eps, momentum=momentum), nn.ReLU(inplace=True))
        self.stem1 = nn.Sequential(_conv_holder(64, 64, 3, 1, 1), BatchNorm(64, eps=eps, momentum=momentum))
        self.stem2 = nn.Sequential(nn.ReLU(inplace=True), _conv_holder(64, 128, 3, 2, 1), BatchNorm(128, eps=eps, momentum=momentum))

        na = self.network_arch
        self.cells = nn.ModuleList(self._build_cells(BatchNorm, F_, B))

        self._aspp_mult = {1: 2, 2: 1, 3: 0.5}[na[-1]]
        self.low_level_conv = nn.Sequential(nn.ReLU(), _conv_holder(FB * 2 ** na[low_level_layer], LOW_LEVEL_C, 1),
                                            BatchNorm(LOW_LEVEL_C, eps=eps, momentum=momentum))
        self.aspp = ASPP_train(FB * fm[na[-1]], ASPP_C, BatchNorm, mult=self._aspp_mult)
        self.conv_aspp = nn.ModuleList()
        for c in self.C_index:
            d = na[c] - na[-1]
            if d == -1:
                self.conv_aspp.append(FactorizedReduce(FB * 2 ** na[c], FB * 2 ** na[-1], BatchNorm, eps=eps, momentum=momentum))
            elif d == -2:
                self.conv_aspp.append(DoubleFactorizedReduce(FB * 2 ** na[c], FB * 2 ** na[-1], BatchNorm, eps=eps, momentum=momentum))
            elif d > 0:
                self.conv_aspp.append(ReLUConvBN(FB * 2 ** na[c], FB * 2 ** na[-1], 1, 1, 0, BatchNorm, eps=eps, momentum=momentum, affine=True))
        self._init_weight()

    DENSE = True          # ADD: cells >= 3 read the projected outputs of all earlier cells (ADD.py:205-238)

    def _build_cells(self, BatchNorm, F_, B):
        """reference ADD.py:171-238 (dense wiring); the siblings without dense links override DENSE."""
        na, fm, FB = self.network_arch, {0: 1, 1: 2, 2: 4, 3: 8}, F_ * B
        cells = []
        for i in range(self.num_net):
            level, prev_level, prev_prev_level = na[i], na[i - 1], na[i - 2]
            downup = int(prev_level - level)
            if i == 0:
                downup = int(0 - level)
                cell = Cell(BatchNorm, B, 64, 128, self.cell_arch, na[i], F_ * fm[level], downup, False, self.DENSE)
            elif i == 1:
                cell = Cell(BatchNorm, B, 128, FB * fm[prev_level], self.cell_arch, na[i], F_ * fm[level], downup, False, self.DENSE)
            elif i == 2 or not self.DENSE:
                cell = Cell(BatchNorm, B, FB * fm[prev_prev_level], FB * fm[prev_level], self.cell_arch, na[i],
                            F_ * fm[level], downup, False, self.DENSE)
            else:
                dense_channels = [F_ * fm[s] for s in na[:i - 1]]
                cell = Cell(BatchNorm, B, dense_channels, FB * fm[prev_level], self.cell_arch, na[i],
                            F_ * fm[level], downup, True, i < self.num_net - 2)
            cells.append(cell)
        return cells

    # ---- init / bookkeeping --------------------------------------------------------------------
    def _init_weight(self):
        """kaiming-normal conv weights, BN γ=1 β=0 (reference ADD.py:491-500)."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                torch.nn.init.kaiming_normal_(m.weight)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()
        rt.bump_generation()

    def _prepare(self):
        self.cw_stem0 = ConvWeights(self.stem0[0].weight, self.stem0[1], cin_pad=8)   # 16 B/pixel in bf16: TMA-able
        self._stem0_packed = None
        self.cw_stem1 = ConvWeights(self.stem1[0].weight, self.stem1[1])
        self.cw_stem2 = ConvWeights(self.stem2[1].weight, self.stem2[2])
        self.cw_low = ConvWeights(self.low_level_conv[1].weight, self.low_level_conv[2])

    def set_precision(self, precision: Optional[str]) -> "ADD":
        self.precision = precision
        self._plans.clear()
        return self

    def _aspp_size(self, size, exponent: int):
        s = 2.0 ** (-1 * exponent)
        return (int((float(size[0]) - 1.0) * s + 1.0), int((float(size[1]) - 1.0) * s + 1.0))

    # ---- plan recording ------------------------------------------------------------------------
    def _emit_trunk(self, b: Builder, x_nchw: torch.Tensor, first: int, last: int, st: dict) -> None:
        """Emit stems (if first == 0) and cells first..last inclusive, updating the trunk state `st`
        (mirrors ADD.py:283-308)."""
        self._ensure_prepared()
        if first == 0:
            n, _, H, W = x_nchw.shape
            h1, w1 = (H - 1) // 2 + 1, (W - 1) // 2 + 1
            t0 = b.alloc(n, h1, w1, 64)
            if b.dtype == torch.bfloat16 and rt.tc_available():
                # layout change + bf16 conversion + im2col fused into the tcgen05 stem kernel
                if self._stem0_packed is None:
                    self._stem0_packed = rt.pack_stem_tc(self.cw_stem0)
                b.stem_nchw(x_nchw, t0, self._stem0_packed, self.cw_stem0.bias, RELU_OUT, "ADD.stem0")
            else:
                img = b.alloc(n, H, W, 8)
                b.nchw_to_nhwc(x_nchw, 3, img, "ADD.input")
                b.conv(img, t0, self.cw_stem0, 2, 1, 1, RELU_OUT, "ADD.stem0")
            # stem2's in-place ReLU mutates stem0 (Q7): every reader sees relu(stem1 output)
            stem0 = b.alloc(n, h1, w1, 64)
            b.conv(t0, stem0, self.cw_stem1, 1, 1, 1, RELU_OUT, "ADD.stem1")
            h2, w2 = (h1 - 1) // 2 + 1, (w1 - 1) // 2 + 1
            stem1 = b.alloc(n, h2, w2, 128)
            b.conv(stem0, stem1, self.cw_stem2, 2, 1, 1, 0, "ADD.stem2")
            st.update(two=[stem0, stem1], dense=[], cur=None, low_cat=None, size=(H, W))
        for i in range(first, last + 1):
            cell = self.cells[i]
            cr = self._concat_may_be_relud(i)
            if not self.DENSE:       # Baselin_Model / AutoDeepLab: every cell reads the two previous outputs only
                concat, _ = cell.emit_cell(b, st["two"][0], st["two"][1], cr)
                st["two"] = [st["two"][1], concat]
                st["cur"] = concat
            elif i < 3:
                concat, dense = cell.emit_cell(b, st["two"][0], st["two"][1], cr)
                st["two"] = [st["two"][1], concat]
                st["dense"].append(dense)
                if i == 2:
                    st["cur"] = concat
            elif i < self.num_net - 2:
                concat, dense = cell.emit_cell(b, list(st["dense"][:-1]), st["cur"], cr)
                st["cur"] = concat
                st["dense"].append(dense)
            elif i == self.num_net - 1:
                st["cur"], _ = cell.emit_cell(b, list(st["dense"]), st["cur"], cr)
            else:
                st["cur"], _ = cell.emit_cell(b, list(st["dense"][:-1]), st["cur"], cr)
            if i == self.low_level_layer:
                src = st["two"][1]
                cat = self.decoder.new_cat(b, src.n, src.h, src.w)
                # stored post-ReLU: the decoder's `_conv` (its only reader) starts with ReLU (decoder.py:13)
                b.conv(src, cat.slice(ASPP_C, LOW_LEVEL_C), self.cw_low, 1, 0, 1, RELU_IN | RELU_OUT, "ADD.low_level_conv")
                st["low_cat"] = cat

    def _concat_may_be_relud(self, i: int) -> bool:
        """True when every reader of cell i's concat outside the cell starts with ReLU: not an exit feature (resized /
        returned raw), the next cell does not up-sample it (ADD.py:71-77 interpolates the raw tensor) and no later cell
        reads it through a bilinear resize as its prev_prev input (ADD.py:83-86)."""
        if i in self.C_index or i >= self.num_net - 1:
            return False
        if self.cells[i + 1].downup_sample == 1:
            return False
        reads_as_prev_prev = (i + 2 < self.num_net) if not self.DENSE else (i == 0 and self.num_net > 2)
        # the prev_prev reader resizes whenever the heights differ (down/up-sampling does not round-trip on even
        # sizes), so only a run of same-resolution cells is safe
        if reads_as_prev_prev and (self.cells[i + 1].downup_sample != 0 or self.cells[i + 2].downup_sample != 0):
            return False
        return True

    def _feature(self, st: dict, i: int) -> View:
        return st["cur"] if (i > 2 or not self.DENSE) else st["two"][1]

    def _emit_exit_lowres(self, b: Builder, y: View, st: dict, i: int, aspp_size, conv_aspp_iter: int,
                          resize: bool = True, relu_feature: bool = False) -> View:
        """[resize →] [conv_aspp →] ASPP → decoder convs; returns fp32 low-res logits (ADD.py:316-323)."""
        # ASPP starts with ReLU (aspp_train.py:35) and is the only reader of what the resize / level adapter produce here,
        # so those store relu(.) and the ASPP convs skip their ReLU-on-load pass (it costs shared-memory bandwidth on
        # the critical path of the 3x3s: ncu r01y, tensor pipe 61 % busy)
        relud = False
        if resize and (y.h < aspp_size[0] or y.w < aspp_size[1]):
            r = b.scratch(y.n, aspp_size[0], aspp_size[1], y.c)
            # EDM's in-place ReLU (Q4) makes the exit interpolate relu(y)
            last_producer = self.network_arch[i] == self.network_arch[-1]
            b.bilinear(y, r, (RELU_IN if relu_feature else 0) | (RELU_OUT if last_producer else 0), "ADD.exit_resize")
            y, relud = r, last_producer
        if self.network_arch[i] != self.network_arch[-1]:
            mod = self.conv_aspp[conv_aspp_iter]
            n_, c_, h_, w_ = mod.out_shape(y.n, y.c, y.h, y.w)
            a_in = b.scratch(n_, h_, w_, c_)
            mod.emit(b, y, a_in, RELU_OUT)
            y, relud = a_in, True
        a = b.scratch(y.n, y.h, y.w, ASPP_C)
        self.aspp.emit(b, y, a, IN_RELUD if relud else 0)
        logits = self.decoder.emit_lowres(b, a, st["low_cat"])
        b.release(a)
        return logits

    def _get_plan(self, x: torch.Tensor, kind: str) -> "_NetPlan":
        prec = self.precision or rt.default_precision()
        key = (kind, tuple(x.shape), str(x.device), prec, bool(self.use_cuda_graph))
        p = self._plans.get(key)
        if p is None or p.generation != rt.generation():
            p = _NetPlan(self, tuple(x.shape), x.device, prec, kind)
            self._plans[key] = p
        return p

    # ---- public API (reference signatures) --------------------------------------------------
    def forward(self, x: torch.Tensor) -> List[torch.Tensor]:
        """ADD.py:277-325: list of C logits tensors [N, num_classes, H, W] (fp32).
        In `.train()` (what train.py:227 calls) the dense-wired ADD runs the training-mode forward of `training.py`
        (batch-statistics BatchNorm, autograd through libadd_b200's backward kernels); the sibling wirings have no
        training forward and refuse."""
        if self.training and self.DENSE:
            rt.require_cuda(x)
            from . import training as T
            nc = self._num_classes
            outs = [o[:, :nc] for o in T.add_forward(self, x)]      # channel nc.. of the padded classifier output is padding
            rt.bump_generation()          # running statistics moved: folded eval-mode weights / recorded plans are stale
            return outs
        self._check_eval()
        rt.require_cuda(x)
        plan = self._get_plan(x, "forward")
        plan.set_input(x)
        plan.main.run()
        return plan.fresh_logits()

    def evaluate(self, x: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """Fused eval.py:178-185: forward → per-exit argmax → per-exit confusion matrix, without
        materialising full-resolution logits.  Returns int64 [n_exits, N, nc, nc] (per image)."""
        self._check_eval()
        rt.require_cuda(x)
        plan = self._get_plan(x, "evaluate")
        plan.set_input(x, target)
        plan.main.run()
        return plan.cm

    def get_feature(self, x: torch.Tensor):
        """ADD.py:327-377: (exit-1 logits, raw feature at C_index[0])."""
        self._check_eval()
        rt.require_cuda(x)
        plan = self._get_plan(x, "get_feature")
        plan.set_input(x)
        plan.main.run()
        return plan.fresh_logits()[0], plan.fresh_feature()

    def dynamic_inference(self, x: torch.Tensor, threshold=1.0, confidence='edm', edm=False):
        """ADD.py:379-488 (batch-1 semantics).  Returns (y, earlier_exit, seconds, confidence_value)."""
        self._check_eval()
        rt.require_cuda(x)
        if x.shape[0] != 1:
            raise RuntimeError("dynamic_inference has batch-1 semantics in the reference (ADD.py:421); "
                               "use ADD.dynamic_inference_batch for per-image gating of a batch")
        torch.cuda.synchronize()
        tic = time.perf_counter()
        ys, exits, confs = self.dynamic_inference_batch(x, threshold, confidence, edm)
        torch.cuda.synchronize()
        toc = time.perf_counter()
        return ys[0], int(exits[0]), toc - tic, confs[0]

    def dynamic_inference_batch(self, x: torch.Tensor, threshold=1.0, confidence='edm', edm=False,
                                exit_mode: str = "reference"):
        """Per-image early-exit gating for a batch: each image follows exactly the reference's
        batch-1 control flow; with the EDM gate, exited images are compacted out of the batch so they
        stop consuming later layers.  Returns (list of [1,nc,H,W] logits, list of exit flags, list
        of confidence values).  exit_mode='forward' is a labelled deviation (see dynamic.py)."""
        from .dynamic import run_dynamic
        self._check_eval()
        rt.require_cuda(x)
        return run_dynamic(self, x, threshold, confidence, edm, exit_mode)

    def dynamic_evaluate(self, x: torch.Tensor, target: torch.Tensor, threshold=1.0, edm=False,
                         exit_mode: str = "reference", bind_inputs: bool = False):
        """eval.py:195-221 fused for a batch: EDM-gated early exit → argmax → per-image int64
        confusion matrix [N,nc,nc] (no full-resolution logits are materialised).
        Returns (cm, exit flags, confidence values)."""
        from .dynamic import run_dynamic_evaluate
        self._check_eval()
        rt.require_cuda(x)
        return run_dynamic_evaluate(self, x, target, threshold, edm, exit_mode, bind_inputs)


    def dynamic_evaluate_begin(self, x: torch.Tensor, target: torch.Tensor, threshold=1.0, edm=False,
                               exit_mode: str = "reference", bind_inputs: bool = False):
        """dynamic_evaluate split at the first gate (the host decision): enqueue the trunk, return a handle for
        dynamic_evaluate_finish.  Interleaving begin(batch i+1) before finish(batch i) — on DIFFERENT input buffers with
        bind_inputs=True, e.g. HostPipeline's slots — keeps the GPU busy while the host decides."""
        from .dynamic import begin_dynamic_evaluate
        self._check_eval()
        rt.require_cuda(x)
        return begin_dynamic_evaluate(self, x, target, threshold, edm, exit_mode, bind_inputs)

    def dynamic_evaluate_finish(self, handle):
        from .dynamic import finish_dynamic_evaluate
        return finish_dynamic_evaluate(self, handle)


class _NetPlan:
    """Recorded launch plans + static I/O buffers for one (kind, input shape, precision)."""

    def __init__(self, net: ADD, shape, device, precision: str, kind: str):
        self.generation = rt.generation()
        self.kind = kind
        n, _, H, W = shape
        dtype = rt.act_dtype(precision)
        b = Builder(device, dtype, record=True)
        self.x_static = b.raw(shape, torch.float32)
        self.lowres: List[View] = []       # fp32 decoder-resolution logits per exit (plan buffers)
        self.out_shape = (n, net._num_classes, H, W)
        self.cm = None
        self.feature = None
        nc = net._num_classes
        st: dict = {}
        if kind in ("forward", "evaluate"):
            if kind == "evaluate":
                self.gt_static = b.raw((n, H, W), torch.int64)
                n_exits = len([i for i in range(net.num_net) if i in net.C_index or i == net.num_net - 1])
                self.cm = b.raw((n_exits, n, nc, nc), torch.int64)
            aspp_size = net._aspp_size((H, W), net.network_arch[-1] + 2)   # ADD.py:279-280
            it, done, e = 0, -1, 0
            for i in range(net.num_net):
                if not (i in net.C_index or i == net.num_net - 1):
                    continue
                net._emit_trunk(b, self.x_static, done + 1, i, st)
                done = i
                logits = net._emit_exit_lowres(b, net._feature(st, i), st, i, aspp_size, it)
                if net.network_arch[i] != net.network_arch[-1]:
                    it += 1
                if kind == "forward":
                    self.lowres.append(logits)
                else:
                    b.upsample_argmax(logits, H, W, self.gt_static, None, self.cm[e], None, "ADD.upsample_argmax_cm")
                e += 1
        elif kind == "get_feature":
            aspp_size = net._aspp_size((H, W), net.network_arch[-1])        # ADD.py:329-330
            i = min(net.C_index)        # the reference takes the first exit in ASCENDING layer order (ADD.py:366)
            net._emit_trunk(b, self.x_static, 0, i, st)
            self.feature = net._feature(st, i)
            self.lowres.append(net._emit_exit_lowres(b, self.feature, st, i, aspp_size, 0))
        else:
            raise ValueError(kind)
        self.builder = b
        self.main = Plan(b)
        if net.use_cuda_graph:
            self.main.capture()

    # The reference returns FRESH tensors from forward / get_feature (`o1 = model(a); o2 = model(b)` must not alias).
    # The recorded plan ends at the decoder-resolution logits; the final x8 bilinear (decoder.py:28) is launched eagerly
    # after the replay into newly allocated NCHW fp32 tensors — same kernel count as recording it, no copy.
    def fresh_logits(self) -> List[torch.Tensor]:
        dev = self.x_static.device
        s = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        n, nc, H, W = self.out_shape
        outs = []
        for lg in self.lowres:
            out = torch.empty(self.out_shape, dtype=torch.float32, device=dev)
            d = lg.desc()
            check(lib.add_upsample_logits_nchw(ctypes.byref(d), out.data_ptr(), H, W, s), "ADD.upsample_logits")
            outs.append(out)
        return outs

    def fresh_feature(self) -> torch.Tensor:
        """The raw exit feature as a fresh NCHW fp32 tensor (ADD.py:366-377 returns `x` itself)."""
        f = self.feature
        dev = self.x_static.device
        out = torch.empty((f.n, f.c, f.h, f.w), dtype=torch.float32, device=dev)
        d = f.desc()
        check(lib.add_nhwc_to_nchw(ctypes.byref(d), out.data_ptr(), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
              "ADD.feature_nchw")
        return out

    def set_input(self, x: torch.Tensor, target: Optional[torch.Tensor] = None) -> None:
        self.x_static.copy_(x)        # API-edge copy into the plan's static input buffer
        if target is not None:
            self.gt_static.copy_(target)

    @property
    def n_launches(self) -> int:
        return self.main.n_launches


class EDM(AddModule):
    """Early-decision maker — reference ADD.py:502-525: in-place ReLU → 3×3 s2 conv 400→128 → ReLU
    → GAP → Linear 128-64-32-1.  Three launches: conv (ReLU in/out fused), GAP, MLP."""

    def __init__(self):
        super().__init__()
        self.gap = nn.AdaptiveAvgPool2d(1)
        self.relu = nn.ReLU(inplace=True)
        self.conv = _conv_holder(400, 128, 3, 2, 1)
        self.edm = nn.Sequential(nn.Linear(128, 64), nn.ReLU(inplace=True), nn.Linear(64, 32),
                                 nn.ReLU(inplace=True), nn.Linear(32, 1))

    def _prepare(self):
        self.cw = ConvWeights(self.conv.weight)
        self.mlp = [t.detach().float().contiguous() for t in
                    (self.edm[0].weight, self.edm[0].bias, self.edm[2].weight, self.edm[2].bias,
                     self.edm[4].weight, self.edm[4].bias)]

    def emit_edm(self, b: Builder, y: View) -> torch.Tensor:
        """Returns the fp32 [N] confidence tensor.  Does NOT mutate `y`; callers model the
        reference's in-place ReLU side effect (Q4) with ReLU-on-load flags."""
        self._ensure_prepared()
        h, w = (y.h - 1) // 2 + 1, (y.w - 1) // 2 + 1
        t = b.scratch(y.n, h, w, 128)
        b.conv(y, t, self.cw, 2, 1, 1, RELU_IN | RELU_OUT, "EDM.conv")
        pooled = b.raw((y.n, 128), torch.float32)
        b.gap(t, pooled, 0, "EDM.gap")
        out = b.raw((y.n,), torch.float32)
        b.edm_mlp(pooled, y.n, self.mlp, out, "EDM.mlp")
        b.release(t)
        return out

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """ADD.py:515-525.  Returns [N,1].  Like the reference, ReLU is applied to the caller's
        tensor in place when it is a channels_last tensor we can alias; otherwise the side effect
        is not observable and is skipped."""
        self._check_eval()
        rt.require_cuda(x)
        x = x.squeeze(1) if x.dim() == 5 else x
        dtype = x.dtype if x.dtype in (torch.float32, torch.bfloat16) else torch.float32
        b = Builder(x.device, dtype, record=False)
        xv = rt.as_nhwc_view(x, b, dtype)
        out = self.emit_edm(b, xv)
        return out.view(-1, 1)
