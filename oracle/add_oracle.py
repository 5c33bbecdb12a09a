"""CPU oracle for the ADD dense-segmentation forward path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs (plus the informational ``--impl torch_cuda`` comparator,
which runs these same functions on cuda:0 as the stock-PyTorch baseline) may
import it, and there only as the checker or the timed baseline.  The product path (``auto-dynamic-deeplab_b200/``) never
imports this module and has no CPU fallback.

What it is: a *functional* restatement, in plain fp32 PyTorch on the CPU, of the
algorithm of the reference's hot path (``/root/reference/modeling/ADD.py``,
``operations.py``, ``aspp_train.py``, ``decoder.py``, ``utils/metrics.py``).  It
takes a reference-format ``state_dict`` (same 1998 keys) plus the architecture
description and evaluates the network with ``torch.nn.functional`` primitives —
the same ATen arithmetic the reference's ``nn.Module`` graph dispatches to.

Parity pin: the reference ships no tests or golden vectors, so
the oracle is pinned against the reference *itself*: ``tests/golden/make_golden.py``
imports the unmodified reference from ``/root/reference`` in the build container,
runs it on seeded inputs/weights and commits the outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks every function here against those
fixtures (CPU, ``-m "not gpu"``).

Every function cites the reference ``file:line`` it follows.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

# modeling/genotypes.py:5-14
PRIMITIVES = ['none', 'max_pool_3x3', 'avg_pool_3x3', 'skip_connect',
              'sep_conv_3x3', 'sep_conv_5x5', 'dil_conv_3x3', 'dil_conv_5x5']

# searched_arch/autodeeplab/genotype.npy — the cell genotype every script loads
# (eval.py:43,68; train.py:73,98); rows are [branch_index, primitive_index].
AUTODEEPLAB_GENOTYPE = np.array(
    [[0, 7], [1, 4], [2, 4], [3, 6], [5, 4], [8, 4], [11, 5], [13, 5], [19, 7], [18, 5]],
    dtype=np.int64)

BN_EPS = 1e-5  # ADD.py:29,132; aspp_train.py:9; decoder.py:10


@dataclass
class Arch:
    """Architecture description = the ADD constructor arguments (ADD.py:119-125)."""
    network_arch: Sequence[int]
    C_index: Sequence[int]
    cell_arch: np.ndarray = field(default_factory=lambda: AUTODEEPLAB_GENOTYPE.copy())
    num_classes: int = 19
    F: int = 20
    B: int = 5
    low_level_layer: int = 0

    @staticmethod
    def searched_dense(C: int = 2, F: int = 20) -> "Arch":
        """eval.py:42-57 (`--network searched-dense`)."""
        table = {2: ([1, 2, 2, 2, 3, 2, 2, 1, 1, 1, 1, 2], [5]),
                 3: ([1, 2, 3, 2, 2, 3, 2, 3, 2, 3, 2, 3], [3, 7]),
                 4: ([1, 2, 3, 3, 2, 3, 3, 3, 3, 3, 2, 2], [2, 5, 8])}
        na, ci = table[C]
        return Arch(na, ci, F=F, low_level_layer=0)

    @staticmethod
    def autodeeplab_dense(C: int = 2, F: int = 20) -> "Arch":
        """eval.py:66-84 (`--network autodeeplab-dense`)."""
        ci = {2: [5], 3: [3, 7], 4: [2, 5, 8]}[C]
        return Arch([0, 0, 0, 1, 2, 1, 2, 2, 3, 3, 2, 1], ci, F=F, low_level_layer=2)


# --------------------------------------------------------------------------------------
# primitives
# --------------------------------------------------------------------------------------

_BN_TRAIN = {"on": False, "momentum": 0.1}


class bn_training:
    """`with bn_training(momentum):` — every BatchNorm of the op functions below runs in TRAINING mode on one device:
    batch statistics and an in-place update of the `running_mean` / `running_var` tensors of the state dict
    (F.batch_norm(training=True), the path SynchronizedBatchNorm2d takes when it is not replicated by DataParallel,
    sync_batchnorm/batchnorm.py:50-53).  SURVEY §8f row 1 (training forward of the operators)."""

    def __init__(self, momentum: float = 0.1):
        self.momentum = momentum

    def __enter__(self):
        self.prev = dict(_BN_TRAIN)
        _BN_TRAIN.update(on=True, momentum=self.momentum)
        return self

    def __exit__(self, *exc):
        _BN_TRAIN.update(self.prev)
        return False


def _bn(sd: SD, p: str, x: torch.Tensor, eps: float = BN_EPS) -> torch.Tensor:
    """Eval-mode BatchNorm2d (SynchronizedBatchNorm2d falls through to F.batch_norm in
    eval / single-device mode: sync_batchnorm/batchnorm.py:50-53); batch statistics under `bn_training`."""
    if _BN_TRAIN["on"]:
        return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'],
                            sd.get(p + '.weight'), sd.get(p + '.bias'), True, _BN_TRAIN["momentum"], eps)
    return F.batch_norm(x, sd[p + '.running_mean'], sd[p + '.running_var'],
                        sd.get(p + '.weight'), sd.get(p + '.bias'), False, 0.0, eps)


def _bilinear(x: torch.Tensor, size: Sequence[int]) -> torch.Tensor:
    """F.interpolate(mode='bilinear') with its default align_corners=False
    (ADD.py:76,84,89,317; decoder.py:24,28)."""
    return F.interpolate(x, [int(size[0]), int(size[1])], mode='bilinear', align_corners=False)


def relu_conv_bn(sd: SD, p: str, x: torch.Tensor, stride: int = 1, padding: int = 0) -> torch.Tensor:
    """operations.py:18-29 — ReLU → conv → BN; keys `<p>.op.1.weight`, `<p>.op.2.*`."""
    x = F.conv2d(F.relu(x), sd[p + '.op.1.weight'], None, stride, padding)
    return _bn(sd, p + '.op.2', x)


def dil_conv(sd: SD, p: str, x: torch.Tensor, k: int) -> torch.Tensor:
    """operations.py:14-15,32-43 — ReLU → DENSE k×k conv, dilation 2, pad k-1 (no groups) → BN."""
    pad = {3: 2, 5: 4}[k]
    x = F.conv2d(F.relu(x), sd[p + '.op.1.weight'], None, 1, pad, 2)
    return _bn(sd, p + '.op.2', x)


def sep_conv(sd: SD, p: str, x: torch.Tensor, k: int) -> torch.Tensor:
    """operations.py:12-13,46-62 — (ReLU → depthwise k×k → pointwise 1×1 → BN) twice."""
    pad = k // 2
    C = x.shape[1]
    x = F.conv2d(F.relu(x), sd[p + '.op.1.weight'], None, 1, pad, 1, C)
    x = F.conv2d(x, sd[p + '.op.2.weight'])
    x = _bn(sd, p + '.op.3', x)
    x = F.conv2d(F.relu(x), sd[p + '.op.5.weight'], None, 1, pad, 1, C)
    x = F.conv2d(x, sd[p + '.op.6.weight'])
    return _bn(sd, p + '.op.7', x)


def apply_primitive(sd: SD, p: str, name: str, x: torch.Tensor, stride: int = 1) -> torch.Tensor:
    """operations.py:7-16 — the OPS factory (ADD only ever uses stride 1, ADD.py:61; the parameter-free primitives
    also at stride 2 as the supernets build them, cell_level_search.py:19)."""
    if stride != 1 and name not in ('none', 'avg_pool_3x3', 'max_pool_3x3'):
        raise NotImplementedError("strided convolutional primitives are not on the ADD path")
    if name == 'sep_conv_3x3':
        return sep_conv(sd, p, x, 3)
    if name == 'sep_conv_5x5':
        return sep_conv(sd, p, x, 5)
    if name == 'dil_conv_3x3':
        return dil_conv(sd, p, x, 3)
    if name == 'dil_conv_5x5':
        return dil_conv(sd, p, x, 5)
    if name == 'skip_connect':
        return x
    if name == 'none':                      # Zero, operations.py:74-83
        return x.mul(0.) if stride == 1 else x[:, :, ::stride, ::stride].mul(0.)
    if name == 'avg_pool_3x3':              # operations.py:9
        return F.avg_pool2d(x, 3, stride, 1, count_include_pad=False)
    if name == 'max_pool_3x3':              # operations.py:10
        return F.max_pool2d(x, 3, stride, 1)
    raise KeyError(name)


def mixed_op(sd: SD, p: str, x: torch.Tensor, weights: torch.Tensor, training: bool = True, stride: int = 1) -> torch.Tensor:
    """cell_level_search.py:10-29 — MixedOp: sum_k w_k * op_k(x) over genotypes.PRIMITIVES (ops built with affine=False,
    the two pools followed by BatchNorm(affine=False): keys `<p>._ops.<k>.1.running_*`); `training=False` applies only
    the argmax primitive (:27-28).  BatchNorm mode follows `bn_training` like every op function here."""
    from_names = ['none', 'max_pool_3x3', 'avg_pool_3x3', 'skip_connect', 'sep_conv_3x3', 'sep_conv_5x5',
                  'dil_conv_3x3', 'dil_conv_5x5']            # modeling/genotypes.py:5-14

    def one(k):
        name, q = from_names[k], f'{p}._ops.{k}'
        if 'pool' in name:
            return _bn(sd, q + '.1', apply_primitive(sd, q + '.0', name, x, stride))
        return apply_primitive(sd, q, name, x, stride)
    if not training:
        return one(int(torch.argmax(weights)))
    return sum(w * one(k) for k, w in enumerate(weights))


def factorized_reduce(sd: SD, p: str, x: torch.Tensor, step: int = 2) -> torch.Tensor:
    """operations.py:86-101 (step 2) / :104-119 (DoubleFactorizedReduce, step 4):
    ReLU; conv_1 samples the lattice at offset 0, conv_2 the lattice at offset step/2
    (zero beyond the edge); channel-cat; BN."""
    x = F.relu(x)
    off = step // 2
    y = F.pad(x, (0, off, 0, off))[:, :, off:, off:]
    out = torch.cat([F.conv2d(x, sd[p + '.conv_1.weight'], None, step),
                     F.conv2d(y, sd[p + '.conv_2.weight'], None, step)], dim=1)
    return _bn(sd, p + '.bn', out)


def _search_scale_dimension(dim: int, scale: float) -> int:
    """cell_level_search.py:81-83."""
    return int((float(dim) - 1.0) * scale + 1.0) if dim % 2 else int(dim * scale)


def search_cell_forward(sd: SD, p: str, B: int, s0, s1_down, s1_same, s1_up, n_alphas: torch.Tensor,
                        pre_preprocess_sample_rate: float = 1) -> List[torch.Tensor]:
    """cell_level_search.py:96-155 — the supernet cell: preprocess each present s1 (down: FactorizedReduce, same: 1x1,
    up: bilinear x2 then 1x1), resize + pre_preprocess s0, then for each s1 list B steps of MixedOp edges (ops shared
    between the lists), concat of the last B states per list."""
    size = None
    if s1_down is not None:
        s1_down = factorized_reduce(sd, p + '.preprocess_down', s1_down, 2)
        size = s1_down.shape[2:]
    if s1_same is not None:
        s1_same = relu_conv_bn(sd, p + '.preprocess_same', s1_same)
        size = s1_same.shape[2:]
    if s1_up is not None:
        s1_up = _bilinear(s1_up, (_search_scale_dimension(s1_up.shape[2], 2), _search_scale_dimension(s1_up.shape[3], 2)))
        s1_up = relu_conv_bn(sd, p + '.preprocess_up', s1_up)
        size = s1_up.shape[2:]
    if s0 is not None:
        if s0.shape[2] < size[0] or s0.shape[3] < size[1]:
            s0 = _bilinear(s0, size)
        if pre_preprocess_sample_rate >= 1:
            s0 = relu_conv_bn(sd, p + '.pre_preprocess', s0)
        else:
            s0 = factorized_reduce(sd, p + '.pre_preprocess', s0, 2 if pre_preprocess_sample_rate == 0.5 else 4)
    outs = []
    for s1 in (s1_down, s1_same, s1_up):
        if s1 is None:
            continue
        states = [s0 if s0 is not None else 0, s1]
        offset = 0
        for i in range(B):
            new = []
            for j, h in enumerate(states):
                b = offset + j
                if s0 is None and j == 0:          # `_ops[b] is None` (prev_prev_C == -1, :73-75)
                    continue
                new.append(mixed_op(sd, f'{p}._ops.{b}', h, n_alphas[b]))
            states.append(sum(new))
            offset += len(states) - 1
        outs.append(torch.cat(states[-B:], dim=1))
    return outs


def search_aspp(sd: SD, p: str, x: torch.Tensor, padding: int, dilation: int) -> torch.Tensor:
    """operations.py:122-158 — the search-time ASPP head."""
    x = F.relu(x)
    a = F.relu(_bn(sd, p + '.conv11.1', F.conv2d(x, sd[p + '.conv11.0.weight'])))
    b = F.relu(_bn(sd, p + '.conv33.1', F.conv2d(x, sd[p + '.conv33.0.weight'], None, 1, padding, dilation)))
    g = F.adaptive_avg_pool2d(x, 1)
    g = F.relu(_bn(sd, p + '.conv_p.1', F.conv2d(g, sd[p + '.conv_p.0.weight'])))
    g = F.interpolate(g, size=x.shape[2:], mode='bilinear', align_corners=True)
    y = torch.cat([a, b, g], dim=1)
    y = F.relu(_bn(sd, p + '.concate_conv.1', F.conv2d(y, sd[p + '.concate_conv.0.weight'])))
    return F.conv2d(y, sd[p + '.final_conv.weight'])


# --------------------------------------------------------------------------------------
# blocks
# --------------------------------------------------------------------------------------

def _scale_dimension(dim: int, scale: float) -> int:
    """ADD.py:65-66."""
    return int((float(dim) - 1.0) * scale + 1.0)


def executed_cell_dag(cell_arch: np.ndarray, B: int = 5) -> List[List[Tuple[int, int, int]]]:
    """The DAG ADD.Cell.forward actually executes (ADD.py:97-110): ops are *built* in genotype
    row order (ADD.py:59-62) but *consumed* in ascending-branch order through a running
    `ops_index`, so op k is bound to the k-th smallest selected branch.

    Returns, per step i, a list of (state_index j, ops_index, primitive_index)."""
    selected = set(int(b) for b in cell_arch[:, 0])
    steps = []
    offset, ops_index, n_states = 0, 0, 2
    for _ in range(B):
        edges = []
        for j in range(n_states):
            if offset + j in selected:
                edges.append((j, ops_index, int(cell_arch[ops_index, 1])))
                ops_index += 1
        steps.append(edges)
        offset += n_states
        n_states += 1
    return steps


def cell_forward(sd: SD, p: str, arch: Arch, downup: int, dense_in: bool, dense_out: bool,
                 prev_prev, prev: torch.Tensor):
    """ADD.py:69-116 (Cell.forward)."""
    s1 = prev
    if downup == 1:  # ADD.py:71-77
        s1 = _bilinear(s1, (_scale_dimension(s1.shape[2], 2), _scale_dimension(s1.shape[3], 2)))
    if downup == -1:  # ADD.py:42-43,79
        s1 = factorized_reduce(sd, p + '.preprocess', s1)
    else:
        s1 = relu_conv_bn(sd, p + '.preprocess', s1)
    hw = (s1.shape[2], s1.shape[3])
    if not dense_in:  # ADD.py:83-86
        s0 = prev_prev
        if s0.shape[2] != hw[0]:
            s0 = _bilinear(s0, hw)
        s0 = relu_conv_bn(sd, p + '.pre_preprocess', s0)
    else:  # ADD.py:87-93
        parts = []
        for i, t in enumerate(prev_prev):
            if t.shape[2] != hw[0]:
                t = _bilinear(t, hw)
            parts.append(relu_conv_bn(sd, f'{p}.pre_preprocess.{i}', t))
        s0 = relu_conv_bn(sd, p + '.pre_preprocess_1x1', torch.cat(parts, dim=1))
    states = [s0, s1]
    for edges in executed_cell_dag(arch.cell_arch, arch.B):  # ADD.py:97-110
        new = [apply_primitive(sd, f'{p}._ops.{k}', PRIMITIVES[prim], states[j]) for j, k, prim in edges]
        states.append(sum(new))
    concat = torch.cat(states[-arch.B:], dim=1)  # ADD.py:112
    if dense_out:
        return prev, concat, relu_conv_bn(sd, p + '.dense_process', concat)
    return concat


def aspp_train(sd: SD, p: str, x: torch.Tensor, mult: float) -> torch.Tensor:
    """aspp_train.py:34-61."""
    x = F.relu(x)
    outs = []
    outs.append(F.relu(_bn(sd, p + '.aspp1_bn', F.conv2d(x, sd[p + '.aspp1.weight']))))
    for k, d in ((2, 6), (3, 12), (4, 18)):
        dd = int(d * mult)
        outs.append(F.relu(_bn(sd, f'{p}.aspp{k}_bn', F.conv2d(x, sd[f'{p}.aspp{k}.weight'], None, 1, dd, dd))))
    x5 = F.adaptive_avg_pool2d(x, 1)
    x5 = F.relu(_bn(sd, p + '.aspp5_bn', F.conv2d(x5, sd[p + '.aspp5.weight'])))
    # nn.Upsample(align_corners=True) from a 1×1 source is a pure broadcast (aspp_train.py:54-55)
    outs.append(x5.expand(-1, -1, x.shape[2], x.shape[3]))
    x = F.conv2d(torch.cat(outs, 1), sd[p + '.conv1.weight'])
    return _bn(sd, p + '.bn1', x)


def decoder(sd: SD, p: str, x: torch.Tensor, low_level: torch.Tensor, size) -> torch.Tensor:
    """decoder.py:23-29 — only H is compared when deciding to resize (Q8); classifier has a bias."""
    if x.shape[2] != low_level.shape[2]:
        x = _bilinear(x, low_level.shape[2:])
    x = torch.cat((x, low_level), 1)
    x = F.relu(x)
    x = F.relu(_bn(sd, p + '._conv.2', F.conv2d(x, sd[p + '._conv.1.weight'], None, 1, 1)))
    x = F.relu(_bn(sd, p + '._conv.5', F.conv2d(x, sd[p + '._conv.4.weight'], None, 1, 1)))
    x = F.conv2d(x, sd[p + '._conv.7.weight'], sd[p + '._conv.7.bias'])
    return _bilinear(x, size)


def edm_forward(sd: SD, y: torch.Tensor) -> torch.Tensor:
    """ADD.py:515-525 (EDM.forward).  Pure function: the reference's in-place ReLU side effect on
    its argument (Q4) is modelled by the caller (`add_dynamic_inference`)."""
    x = F.relu(y)
    x = F.relu(F.conv2d(x, sd['conv.weight'], None, 2, 1))
    x = F.adaptive_avg_pool2d(x, 1).view(y.shape[0], -1)
    x = F.relu(F.linear(x, sd['edm.0.weight'], sd['edm.0.bias']))
    x = F.relu(F.linear(x, sd['edm.2.weight'], sd['edm.2.bias']))
    return F.linear(x, sd['edm.4.weight'], sd['edm.4.bias'])


def normalized_shannon_entropy(x: torch.Tensor, num_class: int = 19) -> float:
    """operations.py:161-170 — summed over batch AND pixels, divided by H·W only (Q5)."""
    e = (F.softmax(x, dim=1) * F.log_softmax(x, dim=1)).sum(dim=1)
    e = -(e / math.log(num_class))
    return (e.sum() / (x.shape[2] * x.shape[3])).item()


def confidence_max(x: torch.Tensor, threshold: float) -> float:
    """operations.py:172-180 — fraction of pixels whose max softmax prob exceeds `threshold`."""
    m = F.softmax(x, dim=1).max(1)[0]
    return int((m > threshold).sum()) / (x.shape[2] * x.shape[3])


# --------------------------------------------------------------------------------------
# the network
# --------------------------------------------------------------------------------------

def _cell_kind(arch: Arch, i: int) -> Tuple[int, bool, bool]:
    """(downup_sample, dense_in, dense_out) per ADD.__init__ (ADD.py:171-240)."""
    n = len(arch.network_arch)
    level = arch.network_arch[i]
    downup = int(0 - level) if i == 0 else int(arch.network_arch[i - 1] - level)
    dense_in = i >= 3
    dense_out = i < n - 2 or i < 3
    return downup, dense_in, dense_out


def aspp_mult(arch: Arch) -> float:
    """ADD.py:242-247."""
    return {1: 2, 2: 1, 3: 0.5}[arch.network_arch[-1]]


def _conv_aspp(sd: SD, arch: Arch, it: int, c: int, y: torch.Tensor) -> torch.Tensor:
    """ADD.py:265-273 — level adapter in front of the shared ASPP."""
    d = arch.network_arch[c] - arch.network_arch[-1]
    p = f'conv_aspp.{it}'
    if d == -1:
        return factorized_reduce(sd, p, y, 2)
    if d == -2:
        return factorized_reduce(sd, p, y, 4)
    return relu_conv_bn(sd, p, y)


def _trunk(sd: SD, arch: Arch, x: torch.Tensor, relu_after_exit_feature=None):
    """Generator over the cell stack shared by forward / get_feature / dynamic_inference
    (ADD.py:283-308 ≡ :333-359 ≡ :388-412).  Yields (i, feature, low_level) after every cell; the
    consumer may `.send(True)` to model EDM's in-place ReLU on the exit feature (Q4), which makes
    every later reader of that tensor see relu(feature)."""
    n = len(arch.network_arch)
    stem = F.relu(_bn(sd, 'stem0.1', F.conv2d(x, sd['stem0.0.weight'], None, 2, 1)))
    stem0 = _bn(sd, 'stem1.1', F.conv2d(stem, sd['stem1.0.weight'], None, 1, 1))
    # stem2's in-place ReLU mutates stem0 (Q7) — harmless: every reader re-applies ReLU or... the
    # bilinear reader in cell 0 (ADD.py:84) sees the mutated tensor, so model it explicitly.
    stem0 = F.relu(stem0)
    stem1 = _bn(sd, 'stem2.2', F.conv2d(stem0, sd['stem2.1.weight'], None, 2, 1))
    two = [stem0, stem1]
    dense: List[torch.Tensor] = []
    low_level = None
    cur = None
    for i in range(n):
        downup, dense_in, dense_out = _cell_kind(arch, i)
        p = f'cells.{i}'
        if i < 3:
            two[0], two[1], fm = cell_forward(sd, p, arch, downup, dense_in, dense_out, two[0], two[1])
            dense.append(fm)
            if i == 2:
                cur = two[1]
        elif i < n - 2:
            _, cur, fm = cell_forward(sd, p, arch, downup, dense_in, dense_out, list(dense[:-1]), cur)
            dense.append(fm)
        elif i == n - 1:
            cur = cell_forward(sd, p, arch, downup, dense_in, dense_out, list(dense), cur)
        else:
            cur = cell_forward(sd, p, arch, downup, dense_in, dense_out, list(dense[:-1]), cur)
        if i == arch.low_level_layer:
            low_level = F.relu(two[1])
            low_level = _bn(sd, 'low_level_conv.2', F.conv2d(low_level, sd['low_level_conv.1.weight']))
        feat = cur if i > 2 else two[1]
        mutate = yield i, feat, low_level
        if mutate:
            if i > 2:
                cur = F.relu(cur)
            else:
                two[1] = F.relu(two[1])
                if i == 2:
                    cur = two[1]   # x IS two_last_inputs[1] at i == 2 (ADD.py:399-400): same storage


def add_forward(sd: SD, arch: Arch, x: torch.Tensor) -> List[torch.Tensor]:
    """ADD.py:277-325 — all exits.  aspp_size uses 2^-(L+2) (ADD.py:279-280)."""
    size = (x.shape[2], x.shape[3])
    s = 2.0 ** (-1 * (arch.network_arch[-1] + 2))
    aspp_size = (int((float(size[0]) - 1.0) * s + 1.0), int((float(size[1]) - 1.0) * s + 1.0))
    n = len(arch.network_arch)
    out, it = [], 0
    for i, y, low in _trunk(sd, arch, x):
        if i in arch.C_index or i == n - 1:
            if y.shape[2] < aspp_size[0] or y.shape[3] < aspp_size[1]:
                y = _bilinear(y, aspp_size)
            if arch.network_arch[i] != arch.network_arch[-1]:
                y = _conv_aspp(sd, arch, it, i, y)
                it += 1
            y = aspp_train(sd, 'aspp', y, aspp_mult(arch))
            out.append(decoder(sd, 'decoder', y, low, size))
    return out


def _plain_trunk(sd: SD, arch: Arch, x: torch.Tensor):
    """Cell stack of the non-dense siblings: every cell reads only the two previous outputs
    (baseline_model.py:228-239 ≡ autodeeplab.py:188-197).  Yields (i, feature, low_level) after every cell."""
    n = len(arch.network_arch)
    stem = F.relu(_bn(sd, 'stem0.1', F.conv2d(x, sd['stem0.0.weight'], None, 2, 1)))
    stem0 = F.relu(_bn(sd, 'stem1.1', F.conv2d(stem, sd['stem1.0.weight'], None, 1, 1)))   # stem2's in-place ReLU (Q7)
    stem1 = _bn(sd, 'stem2.2', F.conv2d(stem0, sd['stem2.1.weight'], None, 2, 1))
    two = [stem0, stem1]
    low_level = None
    for i in range(n):
        downup = _cell_kind(arch, i)[0]
        two = [two[1], cell_forward(sd, f'cells.{i}', arch, downup, False, False, two[0], two[1])]
        if i == arch.low_level_layer:
            low_level = _bn(sd, 'low_level_conv.2', F.conv2d(F.relu(two[1]), sd['low_level_conv.1.weight']))
        yield i, two[1], low_level


def baseline_forward(sd: SD, arch: Arch, x: torch.Tensor) -> List[torch.Tensor]:
    """baseline_model.py:224-254 (Baselin_Model.forward): ADD.forward's exits on the non-dense cell stack."""
    size = (x.shape[2], x.shape[3])
    s = 2.0 ** (-1 * (arch.network_arch[-1] + 2))
    aspp_size = (int((float(size[0]) - 1.0) * s + 1.0), int((float(size[1]) - 1.0) * s + 1.0))
    n = len(arch.network_arch)
    out, it = [], 0
    for i, y, low in _plain_trunk(sd, arch, x):
        if i in arch.C_index or i == n - 1:
            if y.shape[2] < aspp_size[0] or y.shape[3] < aspp_size[1]:
                y = _bilinear(y, aspp_size)
            if arch.network_arch[i] != arch.network_arch[-1]:
                y = _conv_aspp(sd, arch, it, i, y)
                it += 1
            y = aspp_train(sd, 'aspp', y, aspp_mult(arch))
            out.append(decoder(sd, 'decoder', y, low, size))
    return out


def autodeeplab_forward(sd: SD, arch: Arch, x: torch.Tensor) -> torch.Tensor:
    """autodeeplab.py:186-204 (AutoDeepLab.forward): one exit after the last cell, the feature goes to ASPP as is."""
    size = (x.shape[2], x.shape[3])
    y = low = None
    for _, y, low in _plain_trunk(sd, arch, x):
        pass
    return decoder(sd, 'decoder', aspp_train(sd, 'aspp', y, aspp_mult(arch)), low, size)


def add_get_feature(sd: SD, arch: Arch, x: torch.Tensor):
    """ADD.py:327-377 — (exit-1 logits, raw feature at C_index[0]); aspp_size uses 2^-L (Q3)."""
    size = (x.shape[2], x.shape[3])
    s = 2.0 ** (-1 * arch.network_arch[-1])
    aspp_size = (int((float(size[0]) - 1.0) * s + 1.0), int((float(size[1]) - 1.0) * s + 1.0))
    for i, y, low in _trunk(sd, arch, x):
        if i in arch.C_index:
            feature = y
            if y.shape[2] < aspp_size[0] or y.shape[3] < aspp_size[1]:
                y = _bilinear(y, aspp_size)
            if arch.network_arch[i] != arch.network_arch[-1]:
                y = _conv_aspp(sd, arch, 0, i, y)
            y = aspp_train(sd, 'aspp', y, aspp_mult(arch))
            return decoder(sd, 'decoder', y, low, size), feature
    return [], []


def add_dynamic_inference(sd: SD, arch: Arch, x: torch.Tensor, threshold: float = 1.0,
                          confidence: str = 'edm', edm_sd: Optional[SD] = None):
    """ADD.py:379-488 for ONE image (batch-1 semantics: tensor truthiness at :421).

    Returns (y, earlier_exit, confidence_value).  'edm' path: ADD.py:394-438, exits when
    edm(y) <= threshold (Q5), EDM's in-place ReLU is visible to the exit and to later cells (Q4),
    aspp_size uses 2^-L (Q3), the last exit never resizes (ADD.py:433-435).
    'entropy'/'max' path: ADD.py:440-488; the reference returns the feature map `x` instead of the
    logits (:488, a bug noted in SURVEY §3.2) — the oracle returns the logits `y` of the exit that
    was taken and documents the deviation."""
    assert x.shape[0] == 1
    size = (x.shape[2], x.shape[3])
    s = 2.0 ** (-1 * arch.network_arch[-1])
    aspp_size = (int((float(size[0]) - 1.0) * s + 1.0), int((float(size[1]) - 1.0) * s + 1.0))
    n = len(arch.network_arch)
    it = 0
    conf_val = None
    gen = _trunk(sd, arch, x)
    msg = None
    while True:
        try:
            i, y, low = gen.send(msg) if msg is not None else next(gen)
        except StopIteration:
            break
        msg = None
        if not (i in arch.C_index or i == n - 1):
            continue
        if confidence == 'edm':
            if i != n - 1:
                conf_val = edm_forward(edm_sd, y)
                y = F.relu(y)  # in-place ReLU inside EDM.forward mutated the feature (Q4)
                msg = True
                if float(conf_val) > threshold:
                    it += 1
                    continue
                if y.shape[2] < aspp_size[0] or y.shape[3] < aspp_size[1]:
                    y = _bilinear(y, aspp_size)
                if arch.network_arch[i] != arch.network_arch[-1]:
                    y = _conv_aspp(sd, arch, it, i, y)
                y = aspp_train(sd, 'aspp', y, aspp_mult(arch))
                return decoder(sd, 'decoder', y, low, size), 1, conf_val
            y = aspp_train(sd, 'aspp', y, aspp_mult(arch))
            return decoder(sd, 'decoder', y, low, size), 0, conf_val
        else:
            # ADD.py:465-470: the exit head is only evaluated when the feature is SMALLER than
            # aspp_size (the aspp/decoder calls are nested inside that `if`).
            if y.shape[2] < aspp_size[0] or y.shape[3] < aspp_size[1]:
                y = _bilinear(y, aspp_size)
                if arch.network_arch[i] != arch.network_arch[-1]:
                    y = _conv_aspp(sd, arch, it, i, y)
                y = aspp_train(sd, 'aspp', y, aspp_mult(arch))
                y = decoder(sd, 'decoder', y, low, size)
            if i != n - 1:
                if confidence == 'entropy':
                    conf_val = normalized_shannon_entropy(y)
                    if conf_val < threshold:
                        return y, 1, conf_val
                else:
                    conf_val = confidence_max(y, threshold)
                    if conf_val > threshold:
                        return y, 1, conf_val
                it += 1
            else:
                return y, 0, conf_val
    raise RuntimeError("unreachable")


# --------------------------------------------------------------------------------------
# Evaluator (integer contract)
# --------------------------------------------------------------------------------------

def generate_matrix(gt: np.ndarray, pred: np.ndarray, num_class: int = 19) -> np.ndarray:
    """utils/metrics.py:34-39 — int64 [num_class, num_class] histogram of num_class*gt+pred over
    pixels with 0 <= gt < num_class (255 = ignore); rows = gt, cols = pred.  Bit-exact contract."""
    gt = np.asarray(gt).reshape(-1).astype(np.int64)
    pred = np.asarray(pred).reshape(-1).astype(np.int64)
    mask = (gt >= 0) & (gt < num_class)
    label = num_class * gt[mask] + pred[mask]
    return np.bincount(label, minlength=num_class ** 2).reshape(num_class, num_class).astype(np.int64)


def mean_iou(cm: np.ndarray) -> float:
    """utils/metrics.py:18-23,49-52 — on the float32 accumulator (Q6), NaN-skipping mean."""
    c = torch.from_numpy(np.asarray(cm)).to(torch.float32)
    iou = torch.diag(c) / (c.sum(1) + c.sum(0) - torch.diag(c))
    num = torch.where(torch.isnan(iou), torch.zeros_like(iou), torch.ones_like(iou)).sum()
    val = torch.where(torch.isnan(iou), torch.zeros_like(iou), iou).sum()
    return (val / num).item()


# --------------------------------------------------------------------------------------
# numpy-only primitives (independent of ATen) for small-case cross checks of the CUDA kernels
# --------------------------------------------------------------------------------------

def np_bilinear_nchw(x: np.ndarray, out_hw: Tuple[int, int]) -> np.ndarray:
    """align_corners=False bilinear, PyTorch's area_pixel_compute_source_index
    in float32: scale=in/out; src=max(0, scale*(dst+0.5)-0.5); i0=floor(src); i1=min(i0+1,in-1)."""
    N, C, H, W = x.shape
    Ho, Wo = out_hw

    def axis(n_in, n_out):
        scale = np.float32(n_in) / np.float32(n_out)
        dst = np.arange(n_out, dtype=np.float32)
        # ATen evaluates scale*(dst+0.5)-0.5 as ONE fused multiply-add (its CPU kernels are built with FP
        # contraction on): emulate the single rounding through float64 (24x24-bit product is exact there).
        fma = (scale.astype(np.float64) * (dst + np.float32(0.5)).astype(np.float64) - 0.5).astype(np.float32)
        src = np.maximum(np.float32(0), fma).astype(np.float32)
        i0 = np.floor(src).astype(np.int64)
        i0 = np.minimum(i0, n_in - 1)
        i1 = np.minimum(i0 + 1, n_in - 1)
        l1 = (src - i0.astype(np.float32)).astype(np.float32)
        return i0, i1, (np.float32(1) - l1).astype(np.float32), l1

    h0, h1, hl0, hl1 = axis(H, Ho)
    w0, w1, wl0, wl1 = axis(W, Wo)
    x = x.astype(np.float32)
    top = x[:, :, h0][:, :, :, w0] * wl0 + x[:, :, h0][:, :, :, w1] * wl1
    bot = x[:, :, h1][:, :, :, w0] * wl0 + x[:, :, h1][:, :, :, w1] * wl1
    return (hl0[None, None, :, None] * top + hl1[None, None, :, None] * bot).astype(np.float32)


def np_conv2d_nchw(x: np.ndarray, w: np.ndarray, stride=1, pad=0, dil=1, groups=1) -> np.ndarray:
    """Direct convolution in float64 accumulate (small cases only)."""
    N, C, H, W = x.shape
    Co, Cg, kh, kw = w.shape
    Ho = (H + 2 * pad - dil * (kh - 1) - 1) // stride + 1
    Wo = (W + 2 * pad - dil * (kw - 1) - 1) // stride + 1
    xp = np.zeros((N, C, H + 2 * pad, W + 2 * pad), np.float64)
    xp[:, :, pad:pad + H, pad:pad + W] = x
    out = np.zeros((N, Co, Ho, Wo), np.float64)
    cpg_out = Co // groups
    for g in range(groups):
        xs = xp[:, g * Cg:(g + 1) * Cg]
        ws = w[g * cpg_out:(g + 1) * cpg_out].astype(np.float64)
        for i in range(kh):
            for j in range(kw):
                patch = xs[:, :, i * dil:i * dil + stride * (Ho - 1) + 1:stride,
                           j * dil:j * dil + stride * (Wo - 1) + 1:stride]
                out[:, g * cpg_out:(g + 1) * cpg_out] += np.einsum('nchw,oc->nohw', patch, ws[:, :, i, j])
    return out.astype(np.float32)


# --------------------------------------------------------------------------------------
# synthetic weights / inputs shared by tests, smoke() and bench.py
# --------------------------------------------------------------------------------------

def sync_batchnorm_train(shards: Sequence[torch.Tensor], weight: Optional[torch.Tensor], bias: Optional[torch.Tensor],
                         running_mean: torch.Tensor, running_var: torch.Tensor, momentum: float = 0.1,
                         eps: float = BN_EPS, sync: bool = True):
    """Training-mode forward of SynchronizedBatchNorm2d over the per-device shards of one step.

    sync=True : modeling/sync_batchnorm/batchnorm.py:55-75 (per-device sum / square-sum, :59-61), :90-111 (reduce over
                the devices, sizes added up) and :113-125 (`_compute_mean_std`: mean = sum/n, sumvar = ssum - sum*mean,
                running stats from the UNBIASED variance, inv_std = clamp(sumvar/n, eps)^-1/2), output :68-75.
    sync=False: the :50-53 path (one device, or DDP where the replication callback never fires — SURVEY §2.2):
                F.batch_norm(training=True) on each shard independently; running stats updated shard after shard.
    Returns (outputs, new_running_mean, new_running_var, mean, inv_std); mean / inv_std are None for sync=False."""
    rm, rv = running_mean.clone(), running_var.clone()
    if not sync:
        outs = [F.batch_norm(x, rm, rv, weight, bias, True, momentum, eps) for x in shards]
        return outs, rm, rv, None, None
    C = shards[0].shape[1]
    flat = [x.reshape(x.shape[0], C, -1) for x in shards]
    size = sum(f.shape[0] * f.shape[2] for f in flat)
    sum_ = sum(f.sum(dim=0).sum(dim=-1) for f in flat)
    ssum = sum((f ** 2).sum(dim=0).sum(dim=-1) for f in flat)
    assert size > 1
    mean = sum_ / size
    sumvar = ssum - sum_ * mean
    unbias_var = sumvar / (size - 1)
    bias_var = sumvar / size
    rm = (1 - momentum) * rm + momentum * mean
    rv = (1 - momentum) * rv + momentum * unbias_var
    inv_std = bias_var.clamp(eps) ** -0.5
    outs = []
    for x, f in zip(shards, flat):
        if weight is not None:
            o = (f - mean.view(1, C, 1)) * (inv_std * weight).view(1, C, 1) + bias.view(1, C, 1)
        else:
            o = (f - mean.view(1, C, 1)) * inv_std.view(1, C, 1)
        outs.append(o.view(x.shape))
    return outs, rm, rv, mean, inv_std


def init_state_dict(arch: Arch, seed: int = 1) -> SD:
    """A random-init, reference-format ``state_dict`` built WITHOUT any model class: the parameter names and shapes
    of ``ADD.__init__`` (ADD.py:119-274; Cell :14-67; operations.py:18-119; aspp_train.py:8-32; decoder.py:8-21) and
    the init rule of ``ADD._init_weight`` (ADD.py:491-500: kaiming-normal conv weights, BN gamma 1 / beta 0, running
    stats (0, 1); the classifier bias keeps nn.Conv2d's default uniform init).  Key set and shapes equal the
    reference's (checked in tests/test_oracle_golden.py); the VALUES are seeded here and are not the reference's RNG
    stream — used where only the workload matters (bench.py's reference arm) and no product module may be imported."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}
    fm = {0: 1, 1: 2, 2: 4, 3: 8}
    F_, B, na = arch.F, arch.B, list(arch.network_arch)
    FB = F_ * B

    def conv(key, co, ci, k):
        fan_in = ci * k * k
        sd[key] = torch.randn(co, ci, k, k, generator=g) * math.sqrt(2.0 / fan_in)      # kaiming_normal_, fan_in, relu gain

    def bn(p, c):
        sd[p + '.weight'] = torch.ones(c)
        sd[p + '.bias'] = torch.zeros(c)
        sd[p + '.running_mean'] = torch.zeros(c)
        sd[p + '.running_var'] = torch.ones(c)
        sd[p + '.num_batches_tracked'] = torch.zeros((), dtype=torch.int64)

    def rcb(p, ci, co, k=1):                     # ReLUConvBN, operations.py:18-29
        conv(p + '.op.1.weight', co, ci, k)
        bn(p + '.op.2', co)

    def fr(p, ci, co):                           # (Double)FactorizedReduce, operations.py:86-119
        conv(p + '.conv_1.weight', co // 2, ci, 1)
        conv(p + '.conv_2.weight', co // 2, ci, 1)
        bn(p + '.bn', co)

    # decoder (decoder.py:8-21) — registered first in ADD.__init__ (ADD.py:150)
    conv('decoder._conv.1.weight', 256, 304, 3); bn('decoder._conv.2', 256)
    conv('decoder._conv.4.weight', 256, 256, 3); bn('decoder._conv.5', 256)
    conv('decoder._conv.7.weight', arch.num_classes, 256, 1)
    bound = 1.0 / math.sqrt(256)
    sd['decoder._conv.7.bias'] = (torch.rand(arch.num_classes, generator=g) * 2 - 1) * bound
    # stems (ADD.py:154-169)
    conv('stem0.0.weight', 64, 3, 3); bn('stem0.1', 64)
    conv('stem1.0.weight', 64, 64, 3); bn('stem1.1', 64)
    conv('stem2.1.weight', 128, 64, 3); bn('stem2.2', 128)
    n = len(na)
    for i in range(n):                           # ADD.py:171-240
        p = f'cells.{i}'
        C = F_ * fm[na[i]]
        downup, dense_in, dense_out = _cell_kind(arch, i)
        prev_C = 128 if i == 0 else FB * fm[na[i - 1]]
        if downup == -1:
            fr(p + '.preprocess', prev_C, C)
        else:
            rcb(p + '.preprocess', prev_C, C)
        for k_, row in enumerate(np.asarray(arch.cell_arch)):
            name = PRIMITIVES[int(row[1])]
            q = f'{p}._ops.{k_}'
            if name.startswith('sep_conv'):
                k = int(name[-1])
                conv(q + '.op.1.weight', C, 1, k); conv(q + '.op.2.weight', C, C, 1); bn(q + '.op.3', C)
                conv(q + '.op.5.weight', C, 1, k); conv(q + '.op.6.weight', C, C, 1); bn(q + '.op.7', C)
            elif name.startswith('dil_conv'):
                k = int(name[-1])
                conv(q + '.op.1.weight', C, C, k); bn(q + '.op.2', C)
        if dense_in:
            chans = [F_ * fm[s] for s in na[:i - 1]]
            for j, c in enumerate(chans):
                rcb(f'{p}.pre_preprocess.{j}', c, C)
            rcb(p + '.pre_preprocess_1x1', len(chans) * C, C)
        else:
            pp_C = 64 if i == 0 else (128 if i == 1 else FB * fm[na[i - 2]])
            rcb(p + '.pre_preprocess', pp_C, C)
        if dense_out:
            rcb(p + '.dense_process', C * B, C)
    conv('low_level_conv.1.weight', 48, FB * 2 ** na[arch.low_level_layer], 1); bn('low_level_conv.2', 48)
    cin = FB * fm[na[-1]]                        # aspp_train.py:8-32
    conv('aspp.aspp1.weight', 256, cin, 1)
    for k_ in (2, 3, 4):
        conv(f'aspp.aspp{k_}.weight', 256, cin, 3)
    conv('aspp.aspp5.weight', 256, cin, 1)
    conv('aspp.conv1.weight', 256, 1280, 1)
    bn('aspp.bn1', 256)
    for k_ in range(1, 6):
        bn(f'aspp.aspp{k_}_bn', 256)
    it = 0
    for c in arch.C_index:                       # ADD.py:265-273
        d = na[c] - na[-1]
        if d in (-1, -2):
            fr(f'conv_aspp.{it}', FB * 2 ** na[c], FB * 2 ** na[-1]); it += 1
        elif d > 0:
            rcb(f'conv_aspp.{it}', FB * 2 ** na[c], FB * 2 ** na[-1]); it += 1
    return sd


def init_edm_state_dict(seed: int = 203) -> SD:
    """EDM parameters (ADD.py:502-513): conv [128,400,3,3] without bias, Linear 128-64-32-1 — nn default
    (kaiming-uniform a=sqrt(5)) init ranges, seeded here (values are not the reference's RNG stream)."""
    g = torch.Generator().manual_seed(seed)

    def uni(shape, fan_in):
        b = 1.0 / math.sqrt(fan_in)
        return (torch.rand(*shape, generator=g) * 2 - 1) * b
    sd = {'conv.weight': uni((128, 400, 3, 3), 400 * 9)}
    for idx, (o, i) in zip((0, 2, 4), ((64, 128), (32, 64), (1, 32))):
        sd[f'edm.{idx}.weight'] = uni((o, i), i)
        sd[f'edm.{idx}.bias'] = uni((o,), i)
    return sd


def randomize_bn_(sd: SD, seed: int = 7) -> SD:
    """Give every BN non-trivial affine params and running stats so BN folding is exercised."""
    g = torch.Generator().manual_seed(seed)
    for k in sorted(sd.keys()):
        if k.endswith('running_mean'):
            sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
        elif k.endswith('running_var'):
            sd[k] = torch.rand(sd[k].shape, generator=g) + 0.5
        elif k.endswith('num_batches_tracked'):
            continue
        else:
            base = k.rsplit('.', 1)[0]
            if base + '.running_mean' in sd:
                if k.endswith('.weight'):
                    sd[k] = torch.rand(sd[k].shape, generator=g) + 0.5
                elif k.endswith('.bias'):
                    sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
    return sd


def calibrate_bn_(sd: SD, arch: Arch, size: Tuple[int, int] = (129, 257), n: int = 8, seed: int = 77,
                  affine_seed: Optional[int] = 23, var_floor: float = 1e-2) -> SD:
    """BN-CALIBRATED weights: one TRAINING-mode forward (batch statistics, momentum 1.0) over a seeded
    synthetic batch writes every BatchNorm's running mean / variance, exactly as the first step of train.py would
    (train.py:227 under model.train()); afterwards the eval-mode activations are O(1) at every depth instead of growing
    to 1e8 with identity statistics, so tolerances on the logits are meaningful.  `affine_seed`:
    also give gamma / beta non-trivial values first (so the BN fold is exercised).  `var_floor`: the image-pool
    branch's BN (aspp5_bn, aspp_train.py:49-53) sees only N values per channel in one step — on a calibration batch of a
    few images its variance can come out at 1e-5..1e-10, an inv_std of ~300 that a running average over thousands of
    training steps never has; running variances are floored at `var_floor` (0 disables).  In place; returns sd."""
    if affine_seed is not None:
        g = torch.Generator().manual_seed(affine_seed)
        for k in sorted(sd.keys()):
            base = k.rsplit('.', 1)[0]
            if base + '.running_mean' in sd:
                if k.endswith('.weight'):
                    sd[k] = torch.rand(sd[k].shape, generator=g) * 0.5 + 0.75
                elif k.endswith('.bias'):
                    sd[k] = torch.randn(sd[k].shape, generator=g) * 0.1
    for k in sd:
        if k.endswith('running_mean') or k.endswith('running_var'):
            sd[k] = sd[k].clone()
    x, _ = synthetic_batch(n, size[0], size[1], seed)
    with torch.no_grad(), bn_training(momentum=1.0):
        add_forward(sd, arch, x)
    if var_floor > 0:
        for k in sd:
            if k.endswith('running_var'):
                sd[k].clamp_(min=var_floor)
    return sd


CITYSCAPES_MEAN = (0.29866842, 0.30135223, 0.30561872)     # dataloaders/datasets/cityscapes.py:53
CITYSCAPES_STD = (0.23925215, 0.23859318, 0.2385942)       # dataloaders/datasets/cityscapes.py:54


def normalize_u8_hwc(img_u8: np.ndarray, mean=CITYSCAPES_MEAN, std=CITYSCAPES_STD) -> np.ndarray:
    """dataloaders/custom_transforms.py:17-24 (Normalize) + :39 (ToTensor's HWC -> CHW): uint8 [N,H,W,3] ->
    float32 [N,3,H,W], with numpy's own type promotion (/= 255.0 stays float32, -= / /= with the float64 mean / std
    tuples go through float64)."""
    img = np.array(img_u8).astype(np.float32)
    img /= 255.0
    img -= mean
    img /= std
    return np.ascontiguousarray(img.transpose((0, 3, 1, 2)))


def synthetic_batch(n: int, h: int, w: int, seed: int = 1234, num_class: int = 19):
    """image ~ N(0,1); labels randint(0,19) with 10 % pixels = 255 (ignore)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, h, w, generator=g)
    gt = torch.randint(0, num_class, (n, h, w), generator=g, dtype=torch.int64)
    ign = torch.rand(n, h, w, generator=g) < 0.1
    gt[ign] = 255
    return x, gt


# --------------------------------------------------------------------------------------
# loader / dump edges
# --------------------------------------------------------------------------------------
CITYSCAPES_VOID = [0, 1, 2, 3, 4, 5, 6, 9, 10, 14, 15, 16, 18, 29, 30, -1]                           # cityscapes.py:44
CITYSCAPES_VALID = [7, 8, 11, 12, 13, 17, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 31, 32, 33]      # cityscapes.py:45


def encode_segmap(mask: np.ndarray, ignore_index: int = 255) -> np.ndarray:
    """dataloaders/datasets/cityscapes.py:85-91 — the in-place cascade, restated literally (on a copy)."""
    mask = mask.copy()
    class_map = dict(zip(CITYSCAPES_VALID, range(len(CITYSCAPES_VALID))))
    for voidc in CITYSCAPES_VOID:
        mask[mask == voidc] = ignore_index
    for validc in CITYSCAPES_VALID:
        mask[mask == validc] = class_map[validc]
    return mask


def full_image_eval_preprocess(img_u8: np.ndarray, mask_u8: np.ndarray, crop_size, mean=CITYSCAPES_MEAN, std=CITYSCAPES_STD):
    """dataloaders/custom_transforms.py:322-347: ToTensor (/255, CHW) -> Normalize -> ZeroPad2d / ConstantPad2d(255) to
    at least crop_size (bottom / right).  img_u8 [H,W,3], mask_u8 [H,W] -> (fp32 [3,Hp,Wp], int64 [Hp,Wp])."""
    t = torch.from_numpy(img_u8).permute(2, 0, 1).contiguous().to(torch.float32).div(255)          # transforms.ToTensor
    m_, s_ = torch.as_tensor(mean, dtype=torch.float32), torch.as_tensor(std, dtype=torch.float32)
    t = (t - m_.view(-1, 1, 1)) / s_.view(-1, 1, 1)                                                  # transforms.Normalize
    mask = torch.from_numpy(mask_u8.astype(np.int64))
    h, w = t.shape[1], t.shape[2]
    pad_tb, pad_lr = max(0, crop_size[0] - h), max(0, crop_size[1] - w)
    t = F.pad(t, (0, pad_lr, 0, pad_tb), value=0.0)
    mask = F.pad(mask, (0, pad_lr, 0, pad_tb), value=255)
    return t, mask


CITYSCAPES_COLOURS = np.array([[128, 64, 128], [244, 35, 232], [70, 70, 70], [102, 102, 156], [190, 153, 153], [153, 153, 153],
                               [250, 170, 30], [220, 220, 0], [107, 142, 35], [152, 251, 152], [0, 130, 180], [220, 20, 60],
                               [255, 0, 0], [0, 0, 142], [0, 0, 70], [0, 60, 100], [0, 80, 100], [0, 0, 230], [119, 11, 32]])


def decode_segmap(label_mask: np.ndarray) -> np.ndarray:
    """dataloaders/utils.py:14-51 (dataset='cityscapes'): float64 [H,W,3] in units of 1/255; labels outside [0,19) keep
    their own value in the three channels."""
    r, g, b = label_mask.copy(), label_mask.copy(), label_mask.copy()
    for ll in range(19):
        r[label_mask == ll] = CITYSCAPES_COLOURS[ll, 0]
        g[label_mask == ll] = CITYSCAPES_COLOURS[ll, 1]
        b[label_mask == ll] = CITYSCAPES_COLOURS[ll, 2]
    rgb = np.zeros((label_mask.shape[0], label_mask.shape[1], 3))
    rgb[:, :, 0], rgb[:, :, 1], rgb[:, :, 2] = r / 255.0, g / 255.0, b / 255.0
    return rgb
