"""CPU: dry run of the launch plans — every plan the GPU paths replay is RECORDED here on a CPU stand-in device
(recording allocates buffers and builds the launch list; it launches nothing), then

 * every recorded launch is handed to its C entry point: without a GPU the call must get past the library's own
   argument / shape / alignment validation (ADD_ERR_BAD_ARG, ADD_ERR_UNSUPPORTED, ADD_ERR_WORKSPACE would be host-side
   rejections) and fail only where the CUDA runtime is first needed (ADD_ERR_CUDA) — so a plan that the host logic
   builds with an operand the kernels do not take is caught here, not on the GPU box;
 * the launch DAG (read / write footprints -> dependencies -> stream schedule) is checked for the properties the
   multi-stream CUDA-graph capture relies on.

No numerics: parity lives in the `-m gpu` tests.  What this pins is the host side of every network variant, precision
and control-flow shape (all exits, get_feature, fused evaluate, early-exit segments / heads with one and three gates)."""
import collections

import pytest
import torch

import util
import add_b200
from add_b200 import dynamic as dyn
from add_b200.ADD import _NetPlan

CPU = torch.device("cpu")
OK_WITHOUT_A_GPU = (0, -3)          # ADD_OK (host-only entry points) / ADD_ERR_CUDA (validation passed, no device to launch on)


@pytest.fixture(autouse=True)
def _no_pinned_memory(monkeypatch):
    # pin_memory() needs a CUDA runtime; the dry run only needs the host buffers to exist
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    yield
    add_b200.runtime.set_tc_enabled(True)


def _validate(builder, what):
    assert builder.launches, what
    codes = collections.Counter()
    for fn, args, tag, meta in builder.launches:
        rc = fn(*args, None)
        codes[rc] += 1
        assert rc in OK_WITHOUT_A_GPU, f"{what}: launch '{tag}' rejected by the library's host-side validation (status {rc})"
        assert meta.get("kernel"), (what, tag)
        assert meta.get("flops", 0) >= 0 and meta.get("bytes", 0) >= 0, (what, tag)
    return codes


def _check_dag(plan, what):
    """Dependencies point backwards, and the stream schedule orders every dependency before its consumer: with vector
    clocks (per launch: the newest launch of every stream known to have completed before it starts — its predecessor on
    its own stream plus the launches whose events it waits for, transitively), each dependency j of launch i must be
    covered by clock(i)[stream of j].  This is exactly what the multi-stream CUDA-graph capture relies on."""
    deps = plan.dependencies()
    n = plan.n_launches
    assert len(deps) == n
    for i, d in enumerate(deps):
        assert all(0 <= j < i for j in d), (what, i, d)
    for n_streams in (1, 2, 4):
        sched = plan.schedule(n_streams)
        assert len(sched) == n
        lane_of = [s for s, _ in sched]
        assert all(0 <= s < n_streams for s in lane_of)
        done_after = []                    # done_after[i][s]: newest launch of stream s complete once launch i has completed
        tail_clock = [[-1] * n_streams for _ in range(n_streams)]       # clock at the tail of every stream
        for i, (s, cross) in enumerate(sched):
            before = list(tail_clock[s])                                 # stream order
            for j in cross:
                assert lane_of[j] != s and j < i, (what, n_streams, i, cross)
                before = [max(a, b) for a, b in zip(before, done_after[j])]
            for j in deps[i]:
                assert before[lane_of[j]] >= j, (what, n_streams, f"launch {i} may start before its dependency {j}")
            after = list(before)
            after[s] = i
            done_after.append(after)
            tail_clock[s] = after


NETS = sorted(util.NET_CASES)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("cname", NETS)
def test_network_plans_record_and_validate(cname, precision):
    """ADD.forward / evaluate / get_feature plans of every network variant (C = 2 / 3 / 4, both genotypes)."""
    spec = util.NET_CASES[cname]
    net = util.make_net(spec)
    h, w = spec["sizes"][0]
    kinds = ["forward", "evaluate"] + (["get_feature"] if net.C_index else [])
    for kind in kinds:
        plan = _NetPlan(net, (2, 3, h, w), CPU, precision, kind)
        codes = _validate(plan.builder, f"{cname}/{precision}/{kind}")
        assert codes[-3] > 0
        _check_dag(plan.main, f"{cname}/{precision}/{kind}")
        if kind == "forward":
            n_exits = len([i for i in range(net.num_net) if i in net.C_index or i == net.num_net - 1])
            assert len(plan.lowres) == n_exits
        kernels = {l[3]["kernel"] for l in plan.builder.launches}
        if precision == "bf16" and "c20" not in cname and spec.get("F", 20) * 2 % 8 == 0:
            assert "conv2d_tc" in kernels, kernels                       # the tensor-core path is what bf16 records


def test_bf16_bench_network_uses_the_tensor_core_kernels_everywhere():
    """searched-dense C=2 F=20 (the benchmarked network) in bf16: every conv-shaped launch is a tcgen05 kernel — no
    CUDA-core conv sneaks into the recorded step (all its channel counts are multiples of 8)."""
    net = add_b200.build_add("searched-dense", 2, 20, seed=1)
    plan = _NetPlan(net, (1, 3, 129, 257), CPU, "bf16", "forward")
    _validate(plan.builder, "bench network")
    kernels = collections.Counter(l[3]["kernel"] for l in plan.builder.launches)
    assert kernels["conv2d_tc"] > 0 and kernels["sepconv_half_tc"] > 0 and kernels["stem_tc"] + kernels.get("stem_conv3x3s2", 0) >= 0
    assert kernels.get("conv2d_ffma", 0) == 0 and kernels.get("sepconv_half", 0) == 0, kernels


@pytest.mark.parametrize("mode", ["logits", "evaluate"])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_early_exit_segments_one_gate(mode, precision):
    """The benchmarked control flow (searched-dense C=2: one EDM gate): trunk segment, early-exit head for the exiting
    images, compacted continuation for the rest — at several image counts, recorded against one runner."""
    spec = util.NET_CASES["searched-dense-C2"]
    net = util.make_net(spec)
    edm = util.make_edm()
    h, w = spec["sizes"][0]
    r = dyn._EdmRunner(net, (4, 3, h, w), CPU, precision, edm, mode, "reference")
    s0 = r.segment(0, 4, None)
    _validate(s0.builder, f"seg0/{mode}/{precision}")
    assert s0.conf is not None and tuple(s0.conf.shape)[0] == 4
    for m_exit in (1, 3):
        hd = r.head(0, m_exit, s0)
        _validate(hd.builder, f"head0 m={m_exit}")
        s1 = r.segment(1, 4 - m_exit, s0)
        _validate(s1.builder, f"seg1 m={4 - m_exit}")
        assert s1.conf is None and s1.out is not None                     # last segment: final head inside the plan
        _check_dag(s1.main, "seg1")
    assert r.segment(1, 3, s0) is r.segment(1, 3, s0)                      # cached by (gate, image count, source segment)
    if mode == "evaluate":
        assert tuple(r.cm_full.shape) == (4, 19, 19) and r.cm_full.dtype == torch.int64


def test_early_exit_segments_three_gates_lineage():
    """Three gated exits (the tests/golden/three_gates.npz network): segments at all four positions and heads at every
    gate record and validate; two continuations with the same (gate, image count) but different source segments are
    different plans (the plan cache is keyed by lineage), the same source gives the cached plan back."""
    net, edm, x, _ = util.make_three_gate_case()
    r = dyn._EdmRunner(net, tuple(x.shape), CPU, "fp32", edm, "evaluate", "reference")
    s0 = r.segment(0, 6, None)
    s1a, s1b = r.segment(1, 6, s0), r.segment(1, 4, s0)
    s2a, s2b = r.segment(2, 3, s1a), r.segment(2, 3, s1b)                 # same (k, m), different lineage
    assert s2a is not s2b and r.segment(2, 3, s1a) is s2a
    s3 = r.segment(3, 2, s2a)
    for name, seg in (("s0", s0), ("s1a", s1a), ("s1b", s1b), ("s2a", s2a), ("s2b", s2b), ("s3", s3)):
        _validate(seg.builder, name)
        _check_dag(seg.main, name)
        if seg.k > 0:
            assert seg.gather is not None and seg.gather.n_launches > 0, name
    assert s3.conf is None and s3.out is not None and s2a.conf is not None
    for k, m, seg in ((0, 2, s0), (1, 3, s1a), (2, 1, s2a), (2, 1, s2b)):
        _validate(r.head(k, m, seg).builder, f"head{k} m={m}")
    assert r.head(2, 1, s2a) is not r.head(2, 1, s2b)
    # every compacted segment gathers from ITS OWN source's state buffers and from nobody else's
    def state_ptrs(seg):
        return {v.buf.data_ptr() for _, v in dyn._state_items(dict(seg.state, two=[None, None]))}
    for seg, src, other in ((s2a, s1a, s1b), (s2b, s1b, s1a)):
        reads = set()
        for fn, args, tag, meta in seg.builder.launches[seg.gather_range[0]:seg.gather_range[1]]:
            reads |= {res[0] for res in meta["reads"]}
        assert reads & state_ptrs(src), "the gather reads nothing of its source segment"
        assert not (reads & (state_ptrs(other) - state_ptrs(src))), "the gather reads another lineage's buffers"


@pytest.mark.parametrize("cname", sorted(util.SIBLING_CASES))
def test_sibling_plans_record_and_validate(cname):
    """Baselin_Model / AutoDeepLab (non-dense wirings over the same kernels)."""
    spec = util.SIBLING_CASES[cname]
    net = util.make_sibling(spec)
    h, w = spec["size"]
    for precision in ("fp32", "bf16"):
        plan = _NetPlan(net, (1, 3, h, w), CPU, precision, "forward")
        _validate(plan.builder, f"{cname}/{precision}")


def test_benchmarked_step_at_full_size():
    """BASELINE config 2 exactly as bench.py runs it (8 x 3 x 1024 x 2048, bf16, uint8 labels, 4 of 8 images exit early):
    the three plans of one step — trunk to the gate, early-exit head for 4 images, continuation for 4 — record at full
    size (buffers are allocated, never touched), every launch validates, and the step is 365 launches of this library's
    kernels whose algorithmic work is what bench.py's `roofline.whole_step` reports (5.40 TFLOP per step, r4h)."""
    net = add_b200.build_add("searched-dense", 2, 20, seed=1)
    torch.manual_seed(203)
    edm = add_b200.EDM().eval()
    r = dyn._EdmRunner(net, (8, 3, 1024, 2048), CPU, "bf16", edm, "evaluate", "reference", label_dtype=torch.uint8)
    s0 = r.segment(0, 8, None)
    h0 = r.head(0, 4, s0)
    s1 = r.segment(1, 4, s0)
    kernels, flops, n = collections.Counter(), 0, 0
    for name, obj in (("trunk", s0), ("early-exit head", h0), ("continuation", s1)):
        _validate(obj.builder, name)
        for fn, args, tag, meta in obj.builder.launches:
            kernels[meta["kernel"]] += 1
            flops += meta["flops"]
            n += 1
    assert n == 365 == s0.n_launches + h0.n_launches + s1.n_launches
    assert kernels["sepconv_half_tc"] == 168 and kernels["conv2d_tc"] == 144 and kernels["upsample_argmax"] == 2
    assert not (set(kernels) & {"conv2d_ffma", "sepconv_half", "depthwise"}), kernels        # no CUDA-core conv in the step
    assert flops == pytest.approx(5.3966e12, rel=1e-3)
    _check_dag(s0.main, "trunk")
    _check_dag(s1.main, "continuation")
    _check_dag(h0.main, "head")


@pytest.mark.parametrize("hw", [(1025, 2049), (769, 769), (513, 1025)])
def test_other_full_sizes_record_and_validate(hw):
    """The authors' real evaluation size (1025 x 2049: every map 2^k + 1), the training crop (769 x 769) and a half-size
    image, bf16, gated: every launch of the three plans passes the library's validation (odd extents, ragged tiles,
    flattened 1x1 tiling, non-multiple-of-8 widths)."""
    net = add_b200.build_add("searched-dense", 2, 20, seed=1)
    torch.manual_seed(203)
    edm = add_b200.EDM().eval()
    r = dyn._EdmRunner(net, (2, 3, *hw), CPU, "bf16", edm, "evaluate", "reference", label_dtype=torch.uint8)
    s0 = r.segment(0, 2, None)
    for name, obj in (("trunk", s0), ("head", r.head(0, 1, s0)), ("continuation", r.segment(1, 1, s0))):
        codes = _validate(obj.builder, f"{hw} {name}")
        assert codes[-3] > 0
    kernels = collections.Counter(l[3]["kernel"] for l in s0.builder.launches)
    assert not (set(kernels) & {"conv2d_ffma", "sepconv_half"}), kernels


def _config5_cases():
    """BASELINE config 5 (tools/microbench.py --sweep): sep_conv / dil_conv 3x3 / 5x5 and ASPP_train at F = 20 / 40 / 80
    across strides 4 / 8 / 16 / 32 (C = F * stride / 4, spatial = 1024 x 2048 / stride; ASPP input = 5C channels)."""
    out = []
    for F in (20, 40, 80):
        for lvl, stride in enumerate((4, 8, 16, 32)):
            C, h, w = F * (1 << lvl), 1024 // stride, 2048 // stride
            for op in ("sep_conv_3x3", "sep_conv_5x5", "dil_conv_3x3", "dil_conv_5x5"):
                out.append((f"F{F}_s{stride}_{op}", "op", (op, C, h, w)))
            out.append((f"F{F}_s{stride}_aspp_in{5 * C}", "aspp", (5 * C, h, w)))
    return out


@pytest.mark.parametrize("name,kind,a", _config5_cases(), ids=[c[0] for c in _config5_cases()])
def test_config5_op_sweep_is_served(name, kind, a):
    """Every shape of BASELINE config 5 records in bf16 and passes the library's validation (one image: the shapes, not
    the batch, decide what the kernels take).  All of them run on the tcgen05 kernels except C = 20 (F = 20 at stride 4:
    40-byte pixel rows cannot be TMA rows — the CUDA-core kernels serve it)."""
    from add_b200.runtime import Builder, View
    BN = torch.nn.BatchNorm2d
    b = Builder(CPU, torch.bfloat16, record=True)
    if kind == "op":
        op, C, h, w = a
        m = add_b200.OPS[op](C, 1, BN, 1e-5, 0.1, True).eval()
        x = View(torch.empty(1, h, w, C, dtype=torch.bfloat16))
        m.emit(b, x, b.alloc(1, h, w, C), 0)
    else:
        cin, h, w = a
        C = cin
        m = add_b200.ASPP_train(cin, 256, BN, mult=1).eval()
        x = View(torch.empty(1, h, w, cin, dtype=torch.bfloat16))
        m.emit(b, x, b.alloc(1, h, w, 256), 0)
    _validate(b, name)
    kernels = {l[3]["kernel"] for l in b.launches}
    cuda_core = kernels & {"conv2d_ffma", "sepconv_half"}
    if C % 8 == 0:
        assert not cuda_core, (name, kernels)
    else:
        assert C in (20, 100) and cuda_core, (name, kernels)      # F = 20 at stride 4 (and its ASPP input 5 * 20)
