"""CPU: host-side logic of the drop-in modules — executed cell DAG, state_dict compatibility,
BN folding, architecture tables."""
import numpy as np
import torch
import torch.nn.functional as F

import util
import add_b200
from add_b200 import runtime as rt
from add_b200.ADD import executed_edges
from util import orc


def test_executed_dag_matches_oracle_and_survey_q1():
    steps = executed_edges(add_b200.AUTODEEPLAB_CELL, 5)
    want = orc.executed_cell_dag(add_b200.AUTODEEPLAB_CELL, 5)
    assert [[(j, k) for j, k, _ in s] for s in want] == steps
    prims = [[add_b200.PRIMITIVES[p] for _, _, p in s] for s in want]
    # s2 = dil5(s0)+sep3(s1); s3 = sep3(s0)+dil3(s1); s4 = sep3(s0)+sep3(s3);
    #            s5 = sep5(s2)+sep5(s4); s6 = dil5(s4)+sep5(s5)
    assert prims == [['dil_conv_5x5', 'sep_conv_3x3'], ['sep_conv_3x3', 'dil_conv_3x3'],
                     ['sep_conv_3x3', 'sep_conv_3x3'], ['sep_conv_5x5', 'sep_conv_5x5'],
                     ['dil_conv_5x5', 'sep_conv_5x5']]
    assert [[j for j, _ in s] for s in steps] == [[0, 1], [0, 1], [0, 3], [2, 4], [4, 5]]


def test_state_dict_has_reference_key_count():
    net = util.make_net(util.NET_CASES["searched-dense-C2"], randomize_bn=False)
    assert len(net.state_dict()) == 1998          # SURVEY §8b [probed]
    keys = set(net.state_dict())
    for k in ("stem0.0.weight", "cells.0.preprocess.conv_1.weight", "cells.3.pre_preprocess.1.op.1.weight",
              "cells.3.pre_preprocess_1x1.op.2.running_var", "cells.5._ops.9.op.6.weight", "aspp.aspp5_bn.bias",
              "decoder._conv.7.bias", "low_level_conv.1.weight"):
        assert k in keys, k
    edm = util.make_edm()
    assert set(edm.state_dict()) == {"conv.weight", "edm.0.weight", "edm.0.bias", "edm.2.weight", "edm.2.bias",
                                     "edm.4.weight", "edm.4.bias"}


def test_bn_fold_equals_batch_norm():
    m, x = util.make_op_case("relu_conv_bn_1x1")
    cw = rt.ConvWeights(m.op[1].weight, m.op[2])
    w = cw.w.permute(3, 2, 0, 1).contiguous()     # back to [Cout,Cin,kh,kw]
    with torch.no_grad():
        y = F.conv2d(F.relu(x), w, cw.bias)
        ref = orc.relu_conv_bn({"m." + k: v for k, v in m.state_dict().items()}, "m", x)
    assert util.rel_err(y, ref) < 1e-5


def test_network_tables_match_reference_drivers():
    assert add_b200.NETWORKS["searched-dense"][2] == ([1, 2, 2, 2, 3, 2, 2, 1, 1, 1, 1, 2], [5], 0)
    assert add_b200.NETWORKS["autodeeplab-dense"][2][2] == 2
    a = orc.Arch.searched_dense(3)
    assert (list(a.network_arch), list(a.C_index)) == tuple(add_b200.NETWORKS["searched-dense"][3][:2])
    assert np.array_equal(orc.AUTODEEPLAB_GENOTYPE, add_b200.AUTODEEPLAB_CELL)


def test_generation_invalidates_on_load():
    m, _ = util.make_op_case("relu_conv_bn_1x1")
    g0 = rt.generation()
    m.load_state_dict(m.state_dict())
    assert rt.generation() > g0


def test_training_mode_is_refused_where_it_is_not_built():
    """The conv modules have a batch-statistics training forward (tests/test_train_forward.py); what has none (ASPP,
    decoder, whole networks and their plans) refuses `.train()` instead of silently running eval-mode arithmetic."""
    import pytest
    import add_b200
    m = add_b200.ASPP_train(40, 32, torch.nn.BatchNorm2d, depth=32).train()
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 40, 4, 4))
    net = util.make_net(util.NET_CASES["searched-dense-C2"]).train()
    with pytest.raises(RuntimeError):          # ADD.forward in .train() is training.add_forward: built, but no CPU path
        net(torch.zeros(1, 3, 33, 65))
    with pytest.raises(NotImplementedError):   # the fused evaluate / gating paths are inference only
        net.evaluate(torch.zeros(1, 3, 33, 65), torch.zeros(1, 33, 65, dtype=torch.int64))
    m2, x = util.make_op_case("dil_conv_3x3_c40")
    with pytest.raises(RuntimeError):          # it has a training forward, but no CPU path exists
        m2.train()(x)


def _fake_plan(specs):
    """specs: [(reads, writes)] with resources (buffer id, channel lo, channel hi)."""
    class _B:
        record = True
        device = None
    b = _B()
    b.launches = [(None, (), f"op{i}", dict(kernel="k", flops=0, bytes=0, reads=list(r), writes=list(w)))
                  for i, (r, w) in enumerate(specs)]
    return rt.Plan(b)


def test_launch_dag_tracks_channel_slices_and_accumulate():
    A, B, CAT, MID = 1, 2, 3, 4
    specs = [
        ([(A, 0, 40)], [(CAT, 0, 40)]),                  # 0: op(s0) -> cat[0:40]
        ([(B, 0, 40)], [(MID, 0, 40)]),                  # 1: sep half1(s1) -> mid
        ([(MID, 0, 40), (CAT, 0, 40)], [(CAT, 0, 40)]),  # 2: sep half2(mid) += cat[0:40]   (RAW on 1, RAW/WAW on 0)
        ([(A, 0, 40)], [(CAT, 40, 80)]),                 # 3: op(s0) -> cat[40:80]          (independent slice)
        ([(CAT, 0, 80)], [(A, 0, 40)]),                  # 4: reads both slices, overwrites s0 (WAR on 0 and 3)
    ]
    plan = _fake_plan(specs)
    assert plan.dependencies() == [[], [], [0, 1], [], [0, 2, 3]]
    sched = plan.schedule(3)
    streams = [s for s, _ in sched]
    assert len({streams[0], streams[1], streams[3]}) == 3          # the three independent chains run side by side
    # every dependency is honoured: same stream (earlier launch) or an explicit cross-stream wait that covers it
    deps = plan.dependencies()
    for i, (si, cross) in enumerate(sched):
        for j in deps[i]:
            sj = streams[j]
            assert sj == si or any(streams[c] == sj and c >= j for c in cross), (i, j)


def test_unknown_footprint_is_a_barrier():
    plan = _fake_plan([([(1, 0, 8)], [(2, 0, 8)]), ([], []), ([(3, 0, 8)], [(4, 0, 8)])])
    deps = plan.dependencies()
    assert deps[1] == [0] and 1 in deps[2]


def test_cout_slice_reads_the_current_bias():
    """Builder.conv runs convs with more than 256 output channels as 256-channel groups of `ConvWeights.cout_slice`.
    `bias` is a public attribute that callers replace after construction (tests/test_gpu_tc.py does): a slice must read
    the CURRENT bias vector, not the (possibly absent) master it was folded from, and must not copy it."""
    from add_b200.runtime import ConvWeights
    w = torch.randn(320, 8, 1, 1, generator=torch.Generator().manual_seed(4))
    cw = ConvWeights(w)                          # no bias at construction
    assert cw.bias is None and cw.cout_slice(256, 64).bias is None
    cw2 = ConvWeights(w)
    cw2.bias = torch.arange(320, dtype=torch.float32)
    s0, s1 = cw2.cout_slice(0, 256), cw2.cout_slice(256, 64)
    assert torch.equal(s0.bias, cw2.bias[:256]) and torch.equal(s1.bias, cw2.bias[256:])
    assert s1.bias.data_ptr() == cw2.bias.data_ptr() + 256 * 4 and s1.bias.data_ptr() % 16 == 0      # a view, 16-byte aligned
    assert torch.equal(s1.w_h, cw2.w_h[..., 256:]) and (s1.cin, s1.cout) == (8, 64)
    bn = torch.nn.BatchNorm2d(320).eval()
    cw3 = ConvWeights(w, bn, bias=torch.ones(320))
    assert torch.allclose(cw3.cout_slice(256, 64).bias, cw3.bias_h[256:])


def test_product_never_imports_the_oracle_or_the_reference():
    """The oracle (oracle/) and the staged reference (baseline/_ref) are CHECKERS: only tests/, smoke() and bench.py's CPU
    legs may load them.  A fresh interpreter that imports the whole product (every submodule of add_b200) must end up with
    neither in sys.modules, and no product source file may name them in an import statement."""
    import subprocess
    import sys
    mods = sorted(p.stem for p in (util.ROOT / "auto-dynamic-deeplab_b200").glob("*.py") if p.stem not in ("__init__", "_build"))
    code = "\n".join([
        "import sys, importlib",
        f"sys.path.insert(0, {str(util.ROOT)!r})",
        "import add_b200",
        f"for m in {mods!r}: importlib.import_module('add_b200.' + m)",
        "bad = [k for k in sys.modules if k == 'oracle' or k.startswith(('oracle.', 'modeling', 'baseline'))]",
        "print('BAD' if bad else 'CLEAN', bad)"])
    out = subprocess.run([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=str(util.ROOT))
    assert out.returncode == 0, out.stderr[-800:]
    assert out.stdout.strip().startswith("CLEAN"), out.stdout
    import re
    for src in sorted((util.ROOT / "auto-dynamic-deeplab_b200").glob("*.py")):
        for ln in src.read_text().splitlines():
            assert not re.match(r"\s*(from|import)\s+(oracle|baseline|modeling)\b", ln), (src.name, ln)


def test_leaving_training_mode_invalidates_folded_weights():
    """An optimiser step mutates parameters in place — torch.optim's too, which knows nothing of this library.  Switching
    a module between .train() and .eval() therefore bumps the generation that folded weights and recorded plans are
    stamped with; repeating the current mode does not."""
    m, _ = util.make_op_case("relu_conv_bn_1x1")
    m.eval()
    g0 = rt.generation()
    m.eval()
    assert rt.generation() == g0
    m.train()
    g1 = rt.generation()
    assert g1 > g0
    with torch.no_grad():
        m.op[1].weight.mul_(2.0)               # what an optimiser step does
    m.eval()
    assert rt.generation() > g1


def test_bench_arms_use_the_same_synthetic_batch_and_network():
    """bench.py's reference arm builds its inputs and network WITHOUT importing the product (its process must not map
    libadd_b200.so): the batch generator it carries must equal the product's and the oracle's (the fp32 feed; the default
    uint8 feed of the b200 arm is an N(0,1) image of the same shape quantised to PNG bytes, `synthetic_batch_u8`), and its hard-coded searched-dense C=2
    path must equal the product's table."""
    import importlib.util
    import sys
    import add_b200
    spec = importlib.util.spec_from_file_location("bench_under_test", util.ROOT / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(bench)
    finally:
        sys.argv = argv
    xb, gb = bench.cpu_synthetic_batch(2, 33, 65, seed=1234)
    xp, gp = add_b200.synthetic_batch(2, 33, 65, seed=1234)
    xo, go = orc.synthetic_batch(2, 33, 65, seed=1234)
    assert torch.equal(xb, xp) and torch.equal(gb, gp) and torch.equal(xb, xo) and torch.equal(gb, go)
    na, ci, low = add_b200.NETWORKS["searched-dense"][2]
    assert (list(na), list(ci), low) == (list(bench.SEARCHED_DENSE_C2[0]), list(bench.SEARCHED_DENSE_C2[1]), bench.SEARCHED_DENSE_C2[2])
    # the two arms describe the same workload
    assert bench.METRIC == "ADD 1024x2048 inference images/sec" and bench.UNIT == "images/s"
