"""CPU: the oracle (oracle/add_oracle.py) against the golden fixtures produced by the UNMODIFIED
reference (tests/golden/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest
import torch

import util
from util import orc

OPS = np.load(util.ROOT / "tests/golden/ops.npz")
NETS = np.load(util.ROOT / "tests/golden/nets.npz")
TOL = 2e-5   # same ATen arithmetic, different op grouping only


def _sd(module, prefix="m"):
    return {f"{prefix}.{k}": v.detach() for k, v in module.state_dict().items()}


def _oracle_op(name, m, x):
    spec = util.OP_CASES[name]
    sd = _sd(m)
    kind, args = spec["kind"], spec["args"]
    if kind == "OPS":
        return orc.apply_primitive(sd, "m", args[0], x, args[2] if len(args) > 2 else 1)
    if kind == "ReLUConvBN":
        return orc.relu_conv_bn(sd, "m", x, args[3], args[4])
    if kind == "FactorizedReduce":
        return orc.factorized_reduce(sd, "m", x, 2)
    if kind == "DoubleFactorizedReduce":
        return orc.factorized_reduce(sd, "m", x, 4)
    raise KeyError(kind)


@pytest.mark.parametrize("name", sorted(util.OP_CASES))
def test_op_matches_reference(name):
    m, x = util.make_op_case(name)
    assert util.weight_checksum(m.state_dict()) == pytest.approx(float(OPS[name + "/wsum"]), rel=1e-12)
    with torch.no_grad():
        y = _oracle_op(name, m, x)
    ref = torch.from_numpy(OPS[name + "/y"])
    assert y.shape == ref.shape
    assert util.rel_err(y, ref) < TOL


def test_aspp_decoder_edm_match_reference():
    m, x = util.make_aspp_case()
    with torch.no_grad():
        y = orc.aspp_train(_sd(m), "m", x, util.ASPP_CASE["mult"])
    assert util.rel_err(y, torch.from_numpy(OPS["aspp/y"])) < TOL
    m, x, low, size = util.make_decoder_case()
    with torch.no_grad():
        y = orc.decoder(_sd(m), "m", x, low, size)
    assert util.rel_err(y, torch.from_numpy(OPS["decoder/y"])) < TOL
    m, x = util.make_edm_case()
    with torch.no_grad():
        y = orc.edm_forward({k: v.detach() for k, v in m.state_dict().items()}, x)
    assert util.rel_err(y, torch.from_numpy(OPS["edm/y"])) < TOL


def test_mixed_cell_matches_reference():
    """Cell with pools / skip_connect / none edges next to the convs (every OPS primitive) vs the reference Cell."""
    m, xpp, xp = util.make_cell_case()
    assert util.weight_checksum(m.state_dict()) == pytest.approx(float(OPS["cell_mixed/wsum"]), rel=1e-12)
    arch = orc.Arch([1], [], util.MIXED_CELL.copy(), 19, 24, 5, 0)
    with torch.no_grad():
        _, concat, dense = orc.cell_forward(_sd(m), "m", arch, 0, False, True, xpp, xp)
    assert util.rel_err(concat, torch.from_numpy(OPS["cell_mixed/concat"])) < TOL
    assert util.rel_err(dense, torch.from_numpy(OPS["cell_mixed/dense"])) < TOL


def test_confidence_scalars_match_reference():
    lg = util.make_logits_case()
    assert orc.normalized_shannon_entropy(lg) == pytest.approx(float(OPS["conf/entropy"]), rel=1e-5)
    assert orc.confidence_max(lg, 0.3) == pytest.approx(float(OPS["conf/max_0.3"]), abs=1e-12)
    assert orc.confidence_max(lg, 0.6) == pytest.approx(float(OPS["conf/max_0.6"]), abs=1e-12)


@pytest.mark.parametrize("cname", sorted(util.make_evaluator_cases()))
def test_evaluator_bit_exact(cname):
    gt, pred = util.make_evaluator_cases()[cname]
    cm = orc.generate_matrix(gt.numpy(), pred.numpy())
    ref = OPS[f"evaluator/{cname}/cm"]
    assert cm.dtype == np.int64 and np.array_equal(cm, ref)
    miou, ref_miou = orc.mean_iou(cm), float(OPS[f"evaluator/{cname}/miou"])
    assert (np.isnan(miou) and np.isnan(ref_miou)) or miou == pytest.approx(ref_miou, rel=1e-6)


@pytest.mark.parametrize("cname", sorted(util.NET_CASES))
def test_network_matches_reference(cname):
    spec = util.NET_CASES[cname]
    net = util.make_net(spec)
    assert util.weight_checksum(net.state_dict()) == pytest.approx(float(NETS[f"{cname}/wsum"]), rel=1e-12)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    arch = util.oracle_arch(spec)
    for (h, w) in spec["sizes"]:
        x, gt = util.make_input(1, h, w)
        tag = f"{cname}/{h}x{w}"
        with torch.no_grad():
            outs = orc.add_forward(sd, arch, x)
        for e, o in enumerate(outs):
            ref = torch.from_numpy(NETS[f"{tag}/forward/{e}"])
            assert o.shape == ref.shape
            assert util.rel_err(o, ref) < 1e-4
            pred = torch.argmax(ref, 1)   # identical predictions -> bit-exact matrix
            assert np.array_equal(orc.generate_matrix(gt.numpy(), pred.numpy()), NETS[f"{tag}/cm/{e}"])
        if spec.get("dynamic"):
            with torch.no_grad():
                lg, feat = orc.add_get_feature(sd, arch, x)
            assert util.rel_err(lg, torch.from_numpy(NETS[f"{tag}/get_feature/logits"])) < 1e-4
            assert feat.double().abs().sum().item() == pytest.approx(float(NETS[f"{tag}/get_feature/feature_sum"]), rel=1e-4)
            edm = util.make_edm()
            edm_sd = {k: v.detach() for k, v in edm.state_dict().items()}
            with torch.no_grad():
                c0 = float(orc.edm_forward(edm_sd, feat))
            assert c0 == pytest.approx(float(NETS[f"{tag}/edm_value"]), rel=1e-3, abs=1e-4)
            for label, thr in (("exit", c0 + 1.0), ("noexit", c0 - 1.0)):
                with torch.no_grad():
                    y, ee, cv = orc.add_dynamic_inference(sd, arch, x, thr, 'edm', edm_sd)
                assert ee == (1 if label == "exit" else 0)
                assert util.rel_err(y, torch.from_numpy(NETS[f"{tag}/dynamic_edm/{label}/y"])) < 1e-4


SIBS = np.load(util.ROOT / "tests/golden/siblings.npz")


@pytest.mark.parametrize("cname", sorted(util.SIBLING_CASES))
def test_sibling_models_match_reference(cname):
    """Baselin_Model / AutoDeepLab (no dense links) restated in the oracle vs the unmodified reference."""
    spec = util.SIBLING_CASES[cname]
    net = util.make_sibling(spec)
    assert util.weight_checksum(net.state_dict()) == pytest.approx(float(SIBS[f"{cname}/wsum"]), rel=1e-12)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    na, ci, low = util.add_b200.NETWORKS[spec["network"]][spec["C"]]
    x, _ = util.make_input(1, *spec["size"])
    with torch.no_grad():
        if spec["cls"] == "Baselin_Model":
            outs = orc.baseline_forward(sd, orc.Arch(na, ci, low_level_layer=low), x)
        else:
            outs = [orc.autodeeplab_forward(sd, orc.Arch(na, [], low_level_layer=low), x)]
    for e, o in enumerate(outs):
        assert util.rel_err(o, torch.from_numpy(SIBS[f"{cname}/forward/{e}"])) < TOL


GATES = np.load(util.ROOT / "tests/golden/gates.npz")
GATE_CASES = [(h, w, conf, label) for (h, w) in util.NET_CASES["searched-dense-C2"]["sizes"]
              for conf in ("entropy", "max") for label in ("exit", "noexit")]


@pytest.mark.parametrize("h,w,conf,label", GATE_CASES)
def test_entropy_and_max_gates_match_reference(h, w, conf, label):
    """ADD.py:440-488 — the 'entropy' and 'max' gates of dynamic_inference, against fixtures produced by the
    unmodified reference (the logits of the exit it took were observed through a hook on its decoder; the reference
    itself returns the feature map `x`, :488 — the oracle returns those logits, a documented deviation)."""
    spec = util.NET_CASES["searched-dense-C2"]
    net = util.make_net(spec)
    assert util.weight_checksum(net.state_dict()) == pytest.approx(float(GATES["searched-dense-C2/wsum"]), rel=1e-12)
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    x, _ = util.make_input(1, h, w)
    k = f"searched-dense-C2/{h}x{w}/{conf}/{label}"
    with torch.no_grad():
        y, ee, cv = orc.add_dynamic_inference(sd, util.oracle_arch(spec), x, float(GATES[k + "/threshold"]), conf)
    assert ee == (1 if label == "exit" else 0)
    assert float(cv) == pytest.approx(float(GATES[k + "/conf"]), rel=1e-4, abs=1e-6)
    assert util.rel_err(y, torch.from_numpy(GATES[k + "/y"])) < TOL


IO = np.load(util.ROOT / "tests/golden/io_edges.npz")


def test_io_edges_oracle_matches_reference():
    """Loader / dump edges: encode_segmap, full_image_eval_preprocess and decode_segmap restated in
    the oracle against fixtures produced by the unmodified reference classes / functions."""
    assert np.array_equal(orc.encode_segmap(np.arange(256, dtype=np.uint8).reshape(16, 16)), IO["encode/all_ids"])
    for name, spec in util.IO_CASES.items():
        img, ids = util.make_io_case(name)
        enc = orc.encode_segmap(ids)
        assert np.array_equal(enc, IO[f"{name}/encoded"])
        t, m = orc.full_image_eval_preprocess(img, enc, spec["crop"])
        assert t.shape == IO[f"{name}/image"].shape and np.array_equal(t.numpy(), IO[f"{name}/image"])    # bit-identical
        assert np.array_equal(m.numpy(), IO[f"{name}/label"])
        assert np.array_equal(orc.decode_segmap(enc.astype(np.int64)), IO[f"{name}/decoded"])


TRAIN = np.load(util.ROOT / "tests/golden/train_step.npz")


def test_train_step_oracle_matches_reference():
    """train.py:216-247 restated with the oracle (its functional graph under `bn_training` + torch autograd + the SGD
    formulas) against the fixture made by the unmodified reference: loss, every parameter's gradient (sum / |sum|, a few
    in full) and every parameter / running statistic after each of two SGD-nesterov steps."""
    spec = util.TRAIN_STEP
    net, x, gt = util.make_train_case()
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    params = [k for k, _ in net.named_parameters()]
    for k in params:
        sd[k].requires_grad_(True)
    bufs = {k: torch.zeros_like(sd[k]) for k in params}
    arch = util.oracle_arch(spec)
    for step in range(spec["steps"]):
        with orc.bn_training(momentum=0.1):
            outs = orc.add_forward(sd, arch, x)
        losses = [torch.nn.functional.cross_entropy(o, gt, ignore_index=255) for o in outs]
        loss = sum(losses) / len(losses)
        grads = torch.autograd.grad(loss, [sd[k] for k in params])
        assert float(loss) == pytest.approx(float(TRAIN[f"step{step}/loss"]), rel=1e-5)
        if step == 0:
            want = {str(k): float(a_) for k, a_ in zip(TRAIN["grad_names"], TRAIN["grad_abs_sum"])}
            assert sorted(want) == sorted(params)
            for k, g_ in zip(params, grads):
                assert float(g_.double().abs().sum()) == pytest.approx(want[k], rel=2e-3, abs=1e-7), k
                if k in util.TRAIN_FULL_GRADS:
                    ref = torch.from_numpy(TRAIN[f"grad/{k}"])
                    assert util.rel_err(g_, ref) < 2e-3, k
        with torch.no_grad():                      # torch.optim.SGD (train.py:126-127): wd, momentum buffer, nesterov
            for k, g_ in zip(params, grads):
                d = g_ + spec["weight_decay"] * sd[k]
                bufs[k] = d.clone() if step == 0 else spec["momentum"] * bufs[k] + d
                d = d + spec["momentum"] * bufs[k]
                sd[k] -= spec["lr"] * d
        # step 0 starts from bit-identical inputs: tight.  The second step's gradient is ill-conditioned (BatchNorm over the
        # 50 samples of the stride-32 level amplifies the 1e-7 differences of the updated parameters: reference against
        # itself-as-oracle differs by up to 1e-1 on single tensors there), so after it only a loose bound holds
        tol = 1e-4 if step == 0 else 5e-2
        for k, a_ in zip(TRAIN[f"step{step}/state_names"], TRAIN[f"step{step}/state_abs_sum"]):
            assert float(sd[str(k)].detach().double().abs().sum()) == pytest.approx(float(a_), rel=tol, abs=1e-6), (step, k)


G3 = np.load(util.ROOT / "tests/golden/three_gates.npz")


@pytest.mark.parametrize("t", range(4))
def test_three_gated_exits_match_reference(t):
    """ADD.py:394-438 with THREE gated exits (C_index = [3, 6, 9]): for every image and threshold of the fixture the oracle
    takes the reference's decision, returns its confidence value (the value of the gate where the image left, or of the
    last gate) and its logits.  Covers exits at the 1st, 2nd and 3rd gate and the run to the final head."""
    net, edm, x, _ = util.make_three_gate_case()
    assert util.weight_checksum(net.state_dict()) == pytest.approx(float(G3["wsum"]), rel=1e-12)
    c = util.THREE_GATES
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    edm_sd = {k: v.detach() for k, v in edm.state_dict().items()}
    arch = orc.Arch(c["network_arch"], c["C_index"], util.cell_arch(), 19, c["F"], c["B"], c["low_level_layer"])
    thr = float(G3["thresholds"][t])
    for i in range(c["n"]):
        with torch.no_grad():
            y, ee, cv = orc.add_dynamic_inference(sd, arch, x[i:i + 1], thr, 'edm', edm_sd)
        k = f"t{t}/img{i}"
        assert ee == int(G3[k + "/exit"]), (t, i)
        assert float(cv) == pytest.approx(float(G3[k + "/conf"]), rel=1e-4, abs=1e-6), (t, i)
        assert float(y.double().abs().sum()) == pytest.approx(float(G3[k + "/y_abs_sum"]), rel=1e-4), (t, i)
        if k + "/y" in G3.files:
            assert util.rel_err(y, torch.from_numpy(G3[k + "/y"])) < TOL, (t, i)
            assert np.array_equal(np.bincount(y.argmax(1).flatten().numpy(), minlength=19), G3[k + "/argmax_hist"]), (t, i)
