"""Shared deterministic test cases (CPU-side construction only; no GPU work here).

Every case is rebuilt from fixed seeds by OUR modules' constructors — the same code path
tests/golden/make_golden.py used when it drove the reference — so no weights are stored."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

import add_b200  # noqa: E402
from oracle import add_oracle as orc  # noqa: E402

BN = torch.nn.BatchNorm2d

# name -> how to build the module and its input
OP_CASES = {
    "sep_conv_3x3_c40": dict(kind="OPS", args=("sep_conv_3x3", 40), x=(2, 40, 13, 17)),
    "sep_conv_5x5_c40": dict(kind="OPS", args=("sep_conv_5x5", 40), x=(2, 40, 13, 17)),
    "sep_conv_3x3_c80": dict(kind="OPS", args=("sep_conv_3x3", 80), x=(1, 80, 9, 20)),
    "sep_conv_5x5_c160": dict(kind="OPS", args=("sep_conv_5x5", 160), x=(1, 160, 6, 7)),
    "sep_conv_3x3_c20": dict(kind="OPS", args=("sep_conv_3x3", 20), x=(1, 20, 8, 8)),
    "dil_conv_3x3_c40": dict(kind="OPS", args=("dil_conv_3x3", 40), x=(2, 40, 13, 17)),
    "dil_conv_5x5_c40": dict(kind="OPS", args=("dil_conv_5x5", 40), x=(2, 40, 13, 17)),
    "dil_conv_5x5_c80": dict(kind="OPS", args=("dil_conv_5x5", 80), x=(1, 80, 9, 11)),
    "dil_conv_3x3_c20": dict(kind="OPS", args=("dil_conv_3x3", 20), x=(1, 20, 7, 9)),
    "avg_pool_3x3_c40": dict(kind="OPS", args=("avg_pool_3x3", 40), x=(2, 40, 13, 17)),
    "max_pool_3x3_c40": dict(kind="OPS", args=("max_pool_3x3", 40), x=(2, 40, 13, 17)),
    "skip_connect_c40": dict(kind="OPS", args=("skip_connect", 40), x=(2, 40, 13, 17)),
    "none_c40": dict(kind="OPS", args=("none", 40), x=(2, 40, 13, 17)),
    "avg_pool_3x3_c24_st2": dict(kind="OPS", args=("avg_pool_3x3", 24, 2), x=(1, 24, 13, 18)),
    "max_pool_3x3_c24_st2": dict(kind="OPS", args=("max_pool_3x3", 24, 2), x=(1, 24, 13, 18)),
    "none_c24_st2": dict(kind="OPS", args=("none", 24, 2), x=(1, 24, 13, 18)),
    "relu_conv_bn_1x1": dict(kind="ReLUConvBN", args=(200, 40, 1, 1, 0), x=(2, 200, 9, 10)),
    "relu_conv_bn_3x3_s2": dict(kind="ReLUConvBN", args=(24, 48, 3, 2, 1), x=(1, 24, 11, 14)),
    "factorized_reduce_even": dict(kind="FactorizedReduce", args=(128, 40), x=(2, 128, 12, 16)),
    "factorized_reduce_odd": dict(kind="FactorizedReduce", args=(64, 80), x=(1, 64, 13, 9)),
    "double_factorized_reduce": dict(kind="DoubleFactorizedReduce", args=(40, 80), x=(1, 40, 14, 19)),
}
ASPP_CASE = dict(C=40, out=32, depth=32, mult=0.5, x=(2, 40, 12, 20))
NET_CASES = {
    "searched-dense-C2": dict(network="searched-dense", C=2, F=20, sizes=[(33, 65), (48, 80)], dynamic=True),
    "autodeeplab-dense-C2": dict(network="autodeeplab-dense", C=2, F=20, sizes=[(33, 65)], dynamic=False),
    "searched-dense-C3": dict(network="searched-dense", C=3, F=20, sizes=[(33, 65)], dynamic=False),
    "searched-dense-C4": dict(network="searched-dense", C=4, F=20, sizes=[(48, 80)], dynamic=False),
}


SIBLING_CASES = {
    "baselin-searched-C2": dict(cls="Baselin_Model", network="searched-dense", C=2, F=20, size=(33, 65)),
    "autodeeplab-net": dict(cls="AutoDeepLab", network="autodeeplab-dense", C=2, F=20, size=(48, 80)),
}


def make_sibling(spec):
    """Baselin_Model / AutoDeepLab with the reference drivers' network paths, seeded, BN-randomised."""
    na, ci, low = add_b200.NETWORKS[spec["network"]][spec["C"]]
    torch.manual_seed(1)
    if spec["cls"] == "Baselin_Model":
        m = add_b200.Baselin_Model(na, ci, add_b200.AUTODEEPLAB_CELL.copy(), 19, add_b200.Args(spec["F"], 5), low)
    else:
        m = add_b200.AutoDeepLab(na, add_b200.AUTODEEPLAB_CELL.copy(), 19, add_b200.Args(spec["F"], 5), low)
    return _randomized(m, 22)


def weight_checksum(sd) -> float:
    return float(sum(v.double().abs().sum().item() for k, v in sorted(sd.items()) if v.dtype.is_floating_point))


def _randomized(module, seed):
    sd = orc.randomize_bn_({k: v.clone() for k, v in module.state_dict().items()}, seed)
    module.load_state_dict(sd, strict=True)
    return module.eval()


def make_op(name):
    spec = OP_CASES[name]
    torch.manual_seed(100 + sorted(OP_CASES).index(name))
    kind, args = spec["kind"], spec["args"]
    if kind == "OPS":
        m = add_b200.OPS[args[0]](args[1], args[2] if len(args) > 2 else 1, BN, 1e-5, 0.1, True)
    else:
        m = getattr(add_b200, kind)(*args, BN)
    return _randomized(m, 11)


def make_op_case(name):
    m = make_op(name)
    g = torch.Generator().manual_seed(500 + sorted(OP_CASES).index(name))
    x = torch.randn(*OP_CASES[name]["x"], generator=g)
    return m, x


def make_aspp_case():
    torch.manual_seed(201)
    m = add_b200.ASPP_train(ASPP_CASE["C"], ASPP_CASE["out"], BN, depth=ASPP_CASE["depth"], mult=ASPP_CASE["mult"])
    m = _randomized(m, 12)
    x = torch.randn(*ASPP_CASE["x"], generator=torch.Generator().manual_seed(601))
    return m, x


def make_decoder_case():
    torch.manual_seed(202)
    m = _randomized(add_b200.Decoder(19, BN), 13)
    g = torch.Generator().manual_seed(602)
    x = torch.randn(1, 256, 5, 7, generator=g)
    low = torch.randn(1, 48, 9, 13, generator=g)
    return m, x, low, (33, 49)


# a cell genotype that uses every primitive (pools, skip_connect, none next to the convs): rows [branch, primitive]
MIXED_CELL = np.array([[0, 1], [1, 4], [2, 2], [4, 3], [5, 6], [8, 0], [9, 5], [12, 2], [14, 7], [19, 1]], dtype=np.int64)
CELL_CASE = dict(prev_prev_C=48, prev_C=64, C_out=24, x_pp=(2, 48, 11, 14), x_p=(2, 64, 11, 14))


def make_cell_case():
    from add_b200.ADD import Cell
    torch.manual_seed(204)
    m = Cell(BN, 5, CELL_CASE["prev_prev_C"], CELL_CASE["prev_C"], MIXED_CELL.copy(), 1, CELL_CASE["C_out"], 0, False, True)
    m = _randomized(m, 14)
    g = torch.Generator().manual_seed(606)
    return m, torch.randn(*CELL_CASE["x_pp"], generator=g), torch.randn(*CELL_CASE["x_p"], generator=g)


def make_edm():
    torch.manual_seed(203)
    return add_b200.EDM().eval()


def make_edm_case():
    m = make_edm()
    x = torch.randn(2, 400, 9, 12, generator=torch.Generator().manual_seed(603))
    return m, x


def make_logits_case():
    return torch.randn(1, 19, 21, 33, generator=torch.Generator().manual_seed(604)) * 3.0


def make_evaluator_cases():
    g = torch.Generator().manual_seed(605)
    cases = {}
    gt = torch.randint(0, 19, (2, 37, 53), generator=g)
    gt[torch.rand(2, 37, 53, generator=g) < 0.1] = 255
    cases["ragged"] = (gt, torch.randint(0, 19, (2, 37, 53), generator=g))
    cases["all_ignored"] = (torch.full((1, 8, 8), 255, dtype=torch.int64), torch.randint(0, 19, (1, 8, 8), generator=g))
    cases["single_pixel"] = (torch.tensor([[[3]]]), torch.tensor([[[5]]]))
    gt = torch.randint(0, 19, (1, 64, 129), generator=g)
    gt[0, :3] = -1
    cases["negative_gt"] = (gt, torch.randint(0, 19, (1, 64, 129), generator=g))
    cases["one_class"] = (torch.full((1, 16, 16), 7, dtype=torch.int64), torch.full((1, 16, 16), 7, dtype=torch.int64))
    return cases


def cell_arch():
    return add_b200.AUTODEEPLAB_CELL.copy()


def net_arch(spec):
    return add_b200.NETWORKS[spec["network"]][spec["C"]]


def make_net(spec, randomize_bn: bool = True):
    m = add_b200.build_add(spec["network"], spec["C"], spec["F"], seed=1)
    return _randomized(m, 21) if randomize_bn else m


def make_input(n, h, w, seed=1234):
    return orc.synthetic_batch(n, h, w, seed)


def oracle_arch(spec) -> "orc.Arch":
    na, ci, low = net_arch(spec)
    return orc.Arch(na, ci, cell_arch(), 19, spec["F"], 5, low)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max|a-b| / max|b| — the parity metric of BASELINE.json (max-norm relative)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


# ---- SynchronizedBatchNorm2d, training mode -----------------------------------
# shards = the per-device inputs of one step (unequal batch sizes on purpose); affine / eps / momentum vary
SYNCBN_CASES = {
    "c8_2shards": dict(C=8, shards=[(2, 5, 7), (3, 5, 7)], eps=1e-5, momentum=0.1, affine=True, seed=31),
    "c40_3shards": dict(C=40, shards=[(1, 9, 6), (2, 9, 6), (1, 9, 6)], eps=1e-5, momentum=0.1, affine=True, seed=32),
    "c16_noaffine": dict(C=16, shards=[(2, 4, 4), (2, 4, 4)], eps=1e-3, momentum=0.3, affine=False, seed=33),
    "c12_tiny_var": dict(C=12, shards=[(2, 3, 5), (1, 3, 5)], eps=1e-2, momentum=0.1, affine=True, seed=34, scale=1e-2),
}


def make_syncbn_case(name):
    """(list of fp32 NCHW shards, BatchNorm state_dict) for SYNCBN_CASES[name]; `c12_tiny_var` has per-channel variance
    below eps, where clamp(var, eps) (synchronised path) and var + eps (F.batch_norm) differ most."""
    spec = SYNCBN_CASES[name]
    g = torch.Generator().manual_seed(spec["seed"])
    C = spec["C"]
    sc = spec.get("scale", 1.0)
    shards = [(torch.randn(n, C, h, w, generator=g) * sc + torch.randn(1, C, 1, 1, generator=g)) for (n, h, w) in spec["shards"]]
    state = {"running_mean": torch.randn(C, generator=g) * 0.1, "running_var": torch.rand(C, generator=g) + 0.5,
             "num_batches_tracked": torch.tensor(0)}
    if spec["affine"]:
        state["weight"] = torch.rand(C, generator=g) + 0.5
        state["bias"] = torch.randn(C, generator=g) * 0.2
    return shards, state


# ---- training-mode forward of the conv operators (batch statistics) -------------------
TRAIN_OP_CASES = ["sep_conv_3x3_c40", "sep_conv_5x5_c40", "sep_conv_3x3_c80", "dil_conv_3x3_c40", "dil_conv_5x5_c40",
                  "dil_conv_5x5_c80", "relu_conv_bn_1x1", "relu_conv_bn_3x3_s2", "factorized_reduce_even",
                  "factorized_reduce_odd", "double_factorized_reduce"]


# ---- MixedOp (cell_level_search.py:10-29) ------------------------------------------------
MIXED_OP_CASES = {
    "c16": dict(C=16, x=(2, 16, 9, 11), seed=41),
    "c40": dict(C=40, x=(2, 40, 13, 17), seed=42),
}


def make_mixed_op_case(name):
    """(our MixedOp with randomised conv weights and running statistics, input, softmax edge weights)."""
    spec = MIXED_OP_CASES[name]
    torch.manual_seed(spec["seed"])
    m = add_b200.MixedOp(spec["C"], 1, BN)
    g = torch.Generator().manual_seed(spec["seed"] + 100)
    with torch.no_grad():
        for k, v in m.state_dict().items():
            if k.endswith("running_mean"):
                v.copy_(torch.randn(v.shape, generator=g) * 0.1)
            elif k.endswith("running_var"):
                v.copy_(torch.rand(v.shape, generator=g) + 0.5)
    x = torch.randn(*spec["x"], generator=g)
    w = torch.softmax(torch.randn(8, generator=g), dim=0)
    return m, x, w


# ---- loader / dump edges ----------------------------------------------------------------------
IO_CASES = {
    "pad_both": dict(h=37, w=64, crop=(40, 70), seed=51),
    "no_pad": dict(h=45, w=80, crop=(40, 70), seed=52),
    "pad_rows": dict(h=33, w=70, crop=(41, 70), seed=53),
}


def make_io_case(name):
    """(uint8 image [H,W,3], uint8 label-id map [H,W] covering every id 0..255 at least once when it fits)."""
    spec = IO_CASES[name]
    g = torch.Generator().manual_seed(spec["seed"])
    img = torch.randint(0, 256, (spec["h"], spec["w"], 3), generator=g, dtype=torch.int64).to(torch.uint8)
    ids = torch.randint(0, 40, (spec["h"], spec["w"]), generator=g, dtype=torch.int64)
    ids.view(-1)[:256] = torch.arange(256)
    return img.numpy(), ids.to(torch.uint8).numpy()


# ---- training step (train.py:216-247) ----------------------------------------------------------------
TRAIN_STEP = dict(network="searched-dense", C=2, F=20, n=2, size=(129, 129), seed=61, lr=0.01, momentum=0.9, weight_decay=4e-5,
                  nesterov=True, steps=2)
# parameters whose full gradients are stored in the fixture (the rest: sum and |sum| per tensor)
TRAIN_FULL_GRADS = ["stem0.0.weight", "stem1.1.weight", "cells.0.preprocess.conv_2.weight", "cells.0._ops.0.op.1.weight",
                    "cells.0._ops.1.op.1.weight", "cells.0._ops.1.op.6.weight", "cells.0._ops.1.op.7.bias",
                    "cells.3.pre_preprocess.1.op.1.weight", "cells.5.preprocess.op.1.weight", "cells.11._ops.6.op.5.weight",
                    "cells.11.pre_preprocess_1x1.op.2.weight", "low_level_conv.1.weight", "aspp.aspp5.weight",
                    "aspp.aspp5_bn.weight", "aspp.aspp3_bn.bias", "decoder._conv.2.weight", "decoder._conv.5.bias",
                    "decoder._conv.7.weight", "decoder._conv.7.bias"]


def make_train_case():
    spec = TRAIN_STEP
    net = make_net(spec)                       # randomised BN affine parameters and running statistics
    g = torch.Generator().manual_seed(spec["seed"])
    x = torch.randn(spec["n"], 3, *spec["size"], generator=g)
    gt = torch.randint(0, 19, (spec["n"], *spec["size"]), generator=g, dtype=torch.int64)
    gt[torch.rand(spec["n"], *spec["size"], generator=g) < 0.1] = 255
    return net, x, gt


def make_op_grad_case(name):
    """(module in train mode, input, seeded cotangent of the output's shape): loss = sum(output * cotangent)."""
    m, x = make_op_case(name)
    m.train()
    g = torch.Generator().manual_seed(700 + sorted(OP_CASES).index(name))
    n, c, h, w = m.out_shape(*x.shape)
    return m, x, torch.randn(n, c, h, w, generator=g)


# ---- supernet cell / search-time ASPP -----------------------------------------------------------------
SEARCH_CELL = dict(B=2, prev_prev_C=16, prev_C_down=None, prev_C_same=24, prev_C_up=32, C_out=16, n=2, h=12, w=16, seed=71)
SEARCH_ASPP = dict(C=24, out=19, pad=6, dil=6, x=(2, 24, 11, 13), seed=72)


def _randomize_running(m, g):
    with torch.no_grad():
        for k, v in m.state_dict().items():
            if k.endswith("running_mean"):
                v.copy_(torch.randn(v.shape, generator=g) * 0.1)
            elif k.endswith("running_var"):
                v.copy_(torch.rand(v.shape, generator=g) + 0.5)
    return m


def make_search_cell_case():
    """(our search Cell, s0, s1_same, s1_up, raw alphas [n_edges, 8], cotangents of the two concats)."""
    c = SEARCH_CELL
    torch.manual_seed(c["seed"])
    m = add_b200.cell_level_search.Cell(c["B"], c["prev_prev_C"], c["prev_C_down"], c["prev_C_same"], c["prev_C_up"], c["C_out"])
    g = torch.Generator().manual_seed(c["seed"] + 100)
    _randomize_running(m, g)
    n, h, w = c["n"], c["h"], c["w"]
    s0 = torch.randn(n, c["prev_prev_C"], h, w, generator=g)
    s1_same = torch.randn(n, c["prev_C_same"], h, w, generator=g)
    s1_up = torch.randn(n, c["prev_C_up"], h // 2, w // 2, generator=g)
    n_edges = sum(2 + i for i in range(c["B"]))
    alphas = torch.randn(n_edges, 8, generator=g) * 0.5
    cots = [torch.randn(n, c["B"] * c["C_out"], h, w, generator=g) for _ in range(2)]
    return m, s0, s1_same, s1_up, alphas, cots


def make_search_aspp_case():
    c = SEARCH_ASPP
    torch.manual_seed(c["seed"])
    m = add_b200.ASPP(c["C"], c["out"], c["pad"], c["dil"])
    g = torch.Generator().manual_seed(c["seed"] + 100)
    _randomize_running(m, g)
    with torch.no_grad():
        for k, v in m.state_dict().items():
            if k.endswith(".1.weight"):
                v.copy_(torch.rand(v.shape, generator=g) + 0.5)
            elif k.endswith(".1.bias"):
                v.copy_(torch.randn(v.shape, generator=g) * 0.1)
    x = torch.randn(*c["x"], generator=g)
    cot = torch.randn(c["x"][0], c["out"], c["x"][2], c["x"][3], generator=g)
    return m, x, cot


# ---- three EDM-gated exits (ADD.py:394-438 with len(C_index) = 3): the case of tests/golden/three_gates.npz -----------
THREE_GATES = dict(network_arch=[1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2], C_index=[3, 6, 9], F=20, B=5, low_level_layer=0,
                   n=6, h=33, w=65, input_seed=77)     # every gated exit at level 2: 400-channel features (EDM, ADD.py:508)


def make_three_gate_case():
    """(our ADD with BN-randomised weights, EDM, x [6,3,33,65], gt) — deterministic (CPU RNG, fixed seeds)."""
    c = THREE_GATES
    torch.manual_seed(1)
    net = add_b200.ADD(c["network_arch"], c["C_index"], cell_arch(), 19, add_b200.Args(c["F"], c["B"]), c["low_level_layer"])
    net = _randomized(net, 21)
    x, gt = make_input(c["n"], c["h"], c["w"], seed=c["input_seed"])
    return net, make_edm(), x, gt


def three_gate_thresholds(G):
    """The thresholds the fixture was made with, in the order the tests walk them (repeats on purpose: plans are reused)."""
    t = [float(v) for v in G["thresholds"]]
    return [t[0], t[1], t[2], t[3], t[2], t[0]]
