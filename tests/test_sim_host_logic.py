"""CPU: the host logic above the launch layer, checked NUMERICALLY without a GPU (tests/sim_backend.py replaces the op
methods of `Builder` by plain-PyTorch closures for the duration of a test; see its header for what that does and does not
cover).  Plans are recorded and replayed exactly as on the GPU — channel-slice concat, accumulate-into-slice node sums,
ReLU-on-store bookkeeping, shared resized features, early-exit segments with host-filled gather indices, result scatter —
and the results are held to the fixtures produced by the unmodified reference (tests/golden/*.npz)."""
import numpy as np
import pytest
import torch

import util
import add_b200
import sim_backend
from util import orc

NETS = np.load(util.ROOT / "tests/golden/nets.npz")
SIBS = np.load(util.ROOT / "tests/golden/siblings.npz")
G3 = np.load(util.ROOT / "tests/golden/three_gates.npz")
TOL = 2e-4          # same ATen arithmetic as the reference up to BN folding and summation order


@pytest.fixture(autouse=True)
def _sim(monkeypatch):
    sim_backend.install(monkeypatch)
    yield


def _agree(a, b):
    return float((a.argmax(1) == b.argmax(1)).float().mean())


@pytest.mark.parametrize("cname", sorted(util.NET_CASES))
def test_forward_all_exits_and_fused_evaluate(cname):
    spec = util.NET_CASES[cname]
    net = util.make_net(spec)
    for (h, w) in spec["sizes"]:
        x, gt = util.make_input(1, h, w)
        tag = f"{cname}/{h}x{w}"
        outs = net(x)
        assert len(outs) == len([k for k in NETS.files if k.startswith(f"{tag}/forward/")])
        for e, o in enumerate(outs):
            ref = torch.from_numpy(NETS[f"{tag}/forward/{e}"])
            assert util.rel_err(o, ref) < TOL, (tag, e)
            assert _agree(o, ref) >= 0.999
        cm = net.evaluate(x, gt)
        for e, o in enumerate(outs):
            assert np.array_equal(cm[e].sum(0).numpy(), orc.generate_matrix(gt.numpy(), o.argmax(1).numpy())), (tag, e)


def test_batch_rows_are_independent():
    spec = util.NET_CASES["searched-dense-C2"]
    net = util.make_net(spec)
    h, w = spec["sizes"][0]
    x, _ = util.make_input(3, h, w, seed=5)
    outs = [o.clone() for o in net(x)]
    for i in range(3):
        single = net(x[i:i + 1])
        for e in range(len(outs)):
            assert util.rel_err(outs[e][i:i + 1], single[e]) < 2e-5          # oneDNN picks batch-dependent blockings


def test_get_feature_and_edm_gate_match_reference():
    cname = "searched-dense-C2"
    spec = util.NET_CASES[cname]
    net = util.make_net(spec)
    edm = util.make_edm()
    for (h, w) in spec["sizes"]:
        tag = f"{cname}/{h}x{w}"
        x, _ = util.make_input(1, h, w)
        lg, feat = net.get_feature(x)
        assert util.rel_err(lg, torch.from_numpy(NETS[f"{tag}/get_feature/logits"])) < TOL
        assert feat.double().abs().sum().item() == pytest.approx(float(NETS[f"{tag}/get_feature/feature_sum"]), rel=1e-4)
        c0 = float(NETS[f"{tag}/edm_value"])
        for label, thr in (("exit", c0 + 1.0), ("noexit", c0 - 1.0)):
            y, ee, _, cv = net.dynamic_inference(x, threshold=thr, confidence='edm', edm=edm)
            assert ee == (1 if label == "exit" else 0)
            assert float(cv) == pytest.approx(float(NETS[f"{tag}/dynamic_edm/{label}/conf"]), rel=1e-3, abs=1e-4)
            assert util.rel_err(y, torch.from_numpy(NETS[f"{tag}/dynamic_edm/{label}/y"])) < TOL


def test_batched_gating_equals_per_image_control_flow():
    """One gate, a batch of 5 with mixed decisions: gathers, compaction and the result scatter of `dynamic_evaluate`."""
    spec = util.NET_CASES["searched-dense-C2"]
    net = util.make_net(spec)
    edm = util.make_edm()
    h, w = spec["sizes"][0]
    x, gt = util.make_input(5, h, w, seed=31)
    vals = [float(net.dynamic_inference(x[i:i + 1], threshold=1e30, confidence='edm', edm=edm)[3]) for i in range(5)]
    for thr in (sorted(vals)[2] - 1e-4 * abs(sorted(vals)[2]) - 1e-6, 1e30, -1e30):
        ref = []
        for i in range(5):
            y1, e1, _, c1 = net.dynamic_inference(x[i:i + 1], threshold=thr, confidence='edm', edm=edm)
            ref.append((y1.clone(), e1, float(c1)))
        ys, flags, confs = net.dynamic_inference_batch(x, thr, 'edm', edm)
        assert flags == [r[1] for r in ref]
        for i in range(5):
            assert util.rel_err(ys[i], ref[i][0]) < 2e-5
        cms, flags2, _ = net.dynamic_evaluate(x, gt, thr, edm)
        assert flags2 == flags
        for i in range(5):
            assert np.array_equal(cms[i].numpy(), orc.generate_matrix(gt[i].numpy(), ref[i][0].argmax(1).numpy())), (thr, i)


def test_three_gated_exits_match_reference_fixture():
    """The body of tests/test_zz_gpu_late.py::test_three_gated_exits_plan_cache_lineage on the CPU stand-in: three gated
    exits, a batch of 6, six calls with different exit patterns on one runner, against the reference fixture."""
    net, edm, x, gt = util.make_three_gate_case()
    n = util.THREE_GATES["n"]
    thr_index = {float(v): t for t, v in enumerate(G3["thresholds"])}
    for thr in util.three_gate_thresholds(G3):
        t = thr_index[thr]
        ys, flags, confs = net.dynamic_inference_batch(x, thr, 'edm', edm)
        ys = [y.clone() for y in ys]
        assert flags == [int(G3[f"t{t}/img{i}/exit"]) for i in range(n)], (t, flags)
        for i in range(n):
            k = f"t{t}/img{i}"
            assert float(confs[i]) == pytest.approx(float(G3[k + "/conf"]), rel=1e-3, abs=1e-4), (t, i)
            assert float(ys[i].double().abs().sum()) == pytest.approx(float(G3[k + "/y_abs_sum"]), rel=1e-3), (t, i)
            if k + "/y" in G3.files:
                want = torch.from_numpy(G3[k + "/y"])
                assert util.rel_err(ys[i], want) < TOL, (t, i)
                assert _agree(ys[i], want) >= 0.999
        for i in range(n):                       # batch-1 control flow of this library
            y1, e1, _, c1 = net.dynamic_inference(x[i:i + 1], threshold=thr, confidence='edm', edm=edm)
            assert e1 == flags[i] and util.rel_err(ys[i], y1) < 2e-5, (t, i)
        cms, flags2, _ = net.dynamic_evaluate(x, gt, thr, edm)
        assert flags2 == flags
        for i in range(n):
            assert np.array_equal(cms[i].numpy(), orc.generate_matrix(gt[i].numpy(), ys[i].argmax(1).numpy())), (thr, i)
    runner = next(v for k, v in net._plans.items() if k[0] == "edm" and k[5] == "logits" and k[1][0] == n)
    assert len({k[0] for k in runner.segments}) >= 3 and len(runner.segments) >= 4, sorted(k[:2] for k in runner.segments)
    assert len({k[0] for k in runner.heads}) >= 2, sorted(k[:2] for k in runner.heads)


@pytest.mark.parametrize("cname", sorted(util.SIBLING_CASES))
def test_sibling_wirings(cname):
    spec = util.SIBLING_CASES[cname]
    net = util.make_sibling(spec)
    x, _ = util.make_input(1, *spec["size"])
    outs = net(x)
    if spec["cls"] == "AutoDeepLab":
        assert outs[0] is None
        outs = [outs[1]]
    for e, o in enumerate(outs):
        assert util.rel_err(o, torch.from_numpy(SIBS[f"{cname}/forward/{e}"])) < TOL


OPS = np.load(util.ROOT / "tests/golden/ops.npz")


@pytest.mark.parametrize("name", sorted(util.OP_CASES))
def test_operator_modules(name):
    """Every operator drop-in called on its own (BN folded on the host, flags, FactorizedReduce's merged 2x2 taps, pools)."""
    m, x = util.make_op_case(name)
    y = m(x)
    ref = torch.from_numpy(OPS[name + "/y"])
    assert tuple(y.shape) == tuple(ref.shape)
    assert util.rel_err(y, ref) < 5e-5


def test_cell_aspp_decoder_edm_modules():
    m, xpp, xp = util.make_cell_case()
    _, concat, dense = m(xpp, xp)
    assert util.rel_err(concat, torch.from_numpy(OPS["cell_mixed/concat"])) < 5e-5
    assert util.rel_err(dense, torch.from_numpy(OPS["cell_mixed/dense"])) < 5e-5
    m, x = util.make_aspp_case()
    assert util.rel_err(m(x), torch.from_numpy(OPS["aspp/y"])) < 5e-5
    m, x, low, size = util.make_decoder_case()
    assert util.rel_err(m(x, low, size), torch.from_numpy(OPS["decoder/y"])) < 5e-5
    m, x = util.make_edm_case()
    assert util.rel_err(m(x), torch.from_numpy(OPS["edm/y"])) < 1e-4


def test_host_pipeline_ragged_final_batch():
    """HostPipeline over batches of 3, 3, 2, 1, 3 images (a loader's final, smaller batch): every batch equals the direct
    call on that batch, nothing is broadcast into a full slot or counted twice, one slot set per batch shape.
    Host images fp32 / labels int64 (the uint8 edges are direct kernel calls the stand-in does not model)."""
    spec = util.NET_CASES["searched-dense-C2"]
    net = util.make_net(spec)
    edm = util.make_edm()
    sizes = [3, 3, 2, 1, 3]
    batches = [util.make_input(n, 33, 65, seed=400 + i) for i, n in enumerate(sizes)]
    _, _, confs = net.dynamic_evaluate(batches[0][0], batches[0][1], -1e30, edm)
    thr = sorted(float(c) for c in confs)[1]
    want = []
    for x, gt in batches:
        cm, flags, _ = net.dynamic_evaluate(x, gt, thr, edm)
        want.append((cm.clone(), list(flags)))
    pipe = add_b200.HostPipeline(net, edm, thr)
    got = [(cm.clone(), list(flags)) for cm, flags in pipe.evaluate(iter(batches))]
    assert [g[0].shape[1] for g in got] == sizes
    for (cm_g, fl_g), (cm_w, fl_w), (x, gt) in zip(got, want, batches):
        assert fl_g == fl_w and torch.equal(cm_g[0], cm_w)
        assert int(cm_g[0].sum()) == int((gt != 255).sum())
    assert len(pipe._slot_sets) == 3
    pipe2 = add_b200.HostPipeline(net)
    for (cm_g, fl), (x, gt) in zip(pipe2.evaluate(iter(batches[1:4])), batches[1:4]):
        assert fl is None and torch.equal(cm_g, net.evaluate(x, gt))
    # the resident-input pipeline over stable buffers gives the same matrices
    rp = add_b200.ResidentPipeline(net, edm, thr)
    bufs = [(batches[0][0].clone(), batches[0][1].clone()), (batches[1][0].clone(), batches[1][1].clone())]
    outs = [(cm.clone(), list(fl)) for cm, fl in rp.evaluate(iter(bufs))]
    assert torch.equal(outs[0][0], want[0][0]) and torch.equal(outs[1][0], want[1][0]) and outs[0][1] == want[0][1]


def test_api_edge_returns_fresh_tensors_and_first_exit_in_layer_order():
    """(1) `forward` / `get_feature` results must not alias plan buffers that the next call overwrites
    (`o1 = model(a); o2 = model(b)`, as flip-TTA does); (2) `get_feature` takes the first exit in ASCENDING layer order
    (ADD.py:366), whatever the order of C_index."""
    spec = util.NET_CASES["searched-dense-C2"]
    net = util.make_net(spec)
    h, w = spec["sizes"][0]
    xa, _ = util.make_input(1, h, w, seed=1)
    xb, _ = util.make_input(1, h, w, seed=2)
    oa = net(xa)
    keep = [o.clone() for o in oa]
    ob = net(xb)
    assert all(torch.equal(a, k) for a, k in zip(oa, keep))                 # the second call did not overwrite the first result
    assert all(a.data_ptr() != b.data_ptr() for a, b in zip(oa, ob)) and not torch.equal(oa[0], ob[0])
    la, fa = net.get_feature(xa)
    la_keep, fa_keep = la.clone(), fa.clone()
    net.get_feature(xb)
    assert torch.equal(la, la_keep) and torch.equal(fa, fa_keep)
    # C_index in non-ascending order
    c = util.THREE_GATES
    torch.manual_seed(1)
    net2 = add_b200.ADD(c["network_arch"], [9, 3, 6], util.cell_arch(), 19, add_b200.Args(c["F"], c["B"]), c["low_level_layer"])
    net2 = util._randomized(net2, 21)
    x, _ = util.make_input(1, 33, 65, seed=3)
    lg, feat = net2.get_feature(x)
    sd = {k: v.detach() for k, v in net2.state_dict().items()}
    with torch.no_grad():
        lg_ref, feat_ref = orc.add_get_feature(sd, orc.Arch(c["network_arch"], [9, 3, 6], util.cell_arch(), 19, c["F"], c["B"],
                                                            c["low_level_layer"]), x)
    assert util.rel_err(lg, lg_ref) < TOL and util.rel_err(feat, feat_ref) < TOL


def test_bf16_error_is_the_error_of_bf16_storage():
    """What the bf16 path's distance from the fp32 reference is made of.  The stand-in run in bf16 is an IDEAL bf16-storage
    implementation: every operator computes in fp32 on bf16-rounded operands and rounds only what it stores (activations,
    the SepConv depthwise tile, weights) — no kernel, no tensor core, no summation-order effects worth the name.  On the
    input and weights of tools/bf16_parity_probe.py (257 x 513, seed 4321) its error against the fp32 oracle must be the
    error MEASURED for the tcgen05 path on the B200 (profiles/r4d_bf16_parity_probe.json, `precision: bf16`): rms within
    0.75x .. 1.35x per exit and weight set, argmax agreement within 3 points.  I.e. the kernels add nothing to what bf16
    storage alone costs on this random-init network."""
    import json
    probe = [r for r in json.loads((util.ROOT / "profiles" / "r4d_bf16_parity_probe.json").read_text())
             if isinstance(r, dict) and r.get("precision") == "bf16" and r.get("size") == "257x513"]
    measured = {(r["weights"], r["exit"]): r for r in probe}
    na, ci, low = add_b200.NETWORKS["searched-dense"][2]
    arch = orc.Arch(na, ci, low_level_layer=low)
    x, _ = orc.synthetic_batch(1, 257, 513, seed=4321)
    for wname in ("randomized_bn", "calibrated_bn"):
        net = add_b200.build_add("searched-dense", 2, 20, seed=1)
        sd = {k: v.clone() for k, v in net.state_dict().items()}
        sd = orc.randomize_bn_(sd, 21) if wname == "randomized_bn" else orc.calibrate_bn_(sd, arch)
        net.load_state_dict(sd)
        net.eval()
        with torch.no_grad():
            ref = orc.add_forward(sd, arch, x)
        net.set_precision("bf16")
        for e, (o, r) in enumerate(zip(net(x), ref)):
            o, r = o.double(), r.double()
            scale = r.abs().max()
            rms = float((o - r).pow(2).mean().sqrt() / scale)
            agree = float((o.argmax(1) == r.argmax(1)).float().mean())
            m = measured[(wname, e)]
            print(f"bf16 storage model [{wname}] exit {e}: rms {rms:.4e} agree {agree:.4f} | measured on the B200: "
                  f"rms {m['rms_rel']:.4e} agree {m['argmax_agree']:.4f}")
            assert 0.75 * m["rms_rel"] <= rms <= 1.35 * m["rms_rel"], (wname, e, rms, m["rms_rel"])
            assert abs(agree - m["argmax_agree"]) <= 0.03, (wname, e, agree, m["argmax_agree"])
