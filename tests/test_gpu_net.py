"""GPU: whole-network parity (all exits, get_feature, EDM-gated dynamic_inference, fused evaluate)
against the golden outputs of the unmodified reference.
Tolerances: fp32 logits <= 1e-3 max-norm relative (north_star), argmax agreement >= 99.9 %,
confusion matrix bit-exact given identical predictions."""
import numpy as np
import pytest
import torch

import util
import add_b200
from util import orc

pytestmark = pytest.mark.gpu
NETS = np.load(util.ROOT / "tests/golden/nets.npz")
DEV = "cuda:0"


def _agree(a, b):
    return float((a.argmax(1) == b.argmax(1)).float().mean())


@pytest.mark.parametrize("cname", sorted(util.NET_CASES))
@pytest.mark.parametrize("graph", [False, True])
def test_forward_all_exits(cname, graph):
    spec = util.NET_CASES[cname]
    net = util.make_net(spec).to(DEV)
    net.use_cuda_graph = graph
    for (h, w) in spec["sizes"]:
        x, gt = util.make_input(1, h, w)
        tag = f"{cname}/{h}x{w}"
        outs = net(x.to(DEV))
        n_exits = len([k for k in NETS.files if k.startswith(f"{tag}/forward/")])
        assert len(outs) == n_exits
        for e, o in enumerate(outs):
            ref = torch.from_numpy(NETS[f"{tag}/forward/{e}"])
            assert tuple(o.shape) == tuple(ref.shape) and o.dtype == torch.float32
            assert util.rel_err(o, ref) < 1e-3
            assert _agree(o.cpu(), ref) >= 0.999
            # Evaluator on OUR predictions vs oracle histogram of the same predictions: bit-exact
            pred = o.argmax(1)
            ev = add_b200.Evaluator(19)
            cm = ev._generate_matrix(gt.to(DEV), pred).cpu().numpy()
            assert np.array_equal(cm, orc.generate_matrix(gt.numpy(), pred.cpu().numpy()))
        # fused head: argmax + confusion matrix without materialising full-res logits
        cms = net.evaluate(x.to(DEV), gt.to(DEV)).cpu().numpy()
        for e, o in enumerate(outs):
            want = orc.generate_matrix(gt.numpy(), o.argmax(1).cpu().numpy())
            assert np.array_equal(cms[e].sum(0), want)


def test_batch_is_per_image():
    """Batch sharding relies on images being independent: batch-of-3 == three batch-of-1 runs."""
    spec = util.NET_CASES["searched-dense-C2"]
    net = util.make_net(spec).to(DEV)
    x, _ = util.make_input(3, 33, 65, seed=4321)
    outs = [o.clone() for o in net(x.to(DEV))]
    for i in range(3):
        single = net(x[i:i + 1].to(DEV))
        for e in range(len(outs)):
            assert util.rel_err(single[e][0], outs[e][i]) < 1e-6


def test_get_feature_and_dynamic_inference():
    cname = "searched-dense-C2"
    spec = util.NET_CASES[cname]
    net = util.make_net(spec).to(DEV)
    edm = util.make_edm().to(DEV)
    for (h, w) in spec["sizes"]:
        tag = f"{cname}/{h}x{w}"
        x, _ = util.make_input(1, h, w)
        xd = x.to(DEV)
        lg, feat = net.get_feature(xd)
        assert util.rel_err(lg, torch.from_numpy(NETS[f"{tag}/get_feature/logits"])) < 1e-3
        assert feat.double().abs().sum().item() == pytest.approx(float(NETS[f"{tag}/get_feature/feature_sum"]), rel=1e-4)
        c0 = float(NETS[f"{tag}/edm_value"])
        got = float(edm(feat.contiguous()))
        assert got == pytest.approx(c0, rel=1e-3, abs=1e-4)
        for label, thr in (("exit", c0 + 1.0), ("noexit", c0 - 1.0)):
            y, ee, secs, cv = net.dynamic_inference(xd, threshold=thr, confidence='edm', edm=edm)
            assert ee == (1 if label == "exit" else 0)
            assert secs > 0
            assert float(cv) == pytest.approx(float(NETS[f"{tag}/dynamic_edm/{label}/conf"]), rel=1e-3, abs=1e-4)
            ref = torch.from_numpy(NETS[f"{tag}/dynamic_edm/{label}/y"])
            assert util.rel_err(y, ref) < 1e-3
            assert _agree(y.cpu(), ref) >= 0.999


def test_entropy_gate_runs_and_matches_oracle():
    spec = util.NET_CASES["searched-dense-C2"]
    net = util.make_net(spec).to(DEV)
    x, _ = util.make_input(1, 33, 65)
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    arch = util.oracle_arch(spec)
    with torch.no_grad():
        y_ref, ee_ref, cv_ref = orc.add_dynamic_inference(sd, arch, x, 0.5, 'entropy')
    y, ee, _, cv = net.dynamic_inference(x.to(DEV), threshold=0.5, confidence='entropy')
    assert ee == ee_ref
    assert cv == pytest.approx(cv_ref, rel=1e-3, abs=1e-4)
    assert util.rel_err(y, y_ref) < 1e-3


GATES = np.load(util.ROOT / "tests/golden/gates.npz")


@pytest.mark.parametrize("conf", ["entropy", "max"])
@pytest.mark.parametrize("label", ["exit", "noexit"])
def test_entropy_and_max_gates_match_reference(conf, label):
    """dynamic_inference(confidence='entropy' | 'max') (ADD.py:440-488) against fixtures made by the unmodified
    reference: same decision, same confidence value, and the logits of the exit taken (the reference returns the
    feature map there, :488; we return the logits — INTEGRATION.md lists the deviation)."""
    spec = util.NET_CASES["searched-dense-C2"]
    net = util.make_net(spec).to(DEV)
    for (h, w) in spec["sizes"]:
        x, _ = util.make_input(1, h, w)
        k = f"searched-dense-C2/{h}x{w}/{conf}/{label}"
        y, ee, secs, cv = net.dynamic_inference(x.to(DEV), threshold=float(GATES[k + "/threshold"]), confidence=conf)
        assert ee == (1 if label == "exit" else 0)
        assert float(cv) == pytest.approx(float(GATES[k + "/conf"]), rel=1e-3, abs=1e-5)
        ref = torch.from_numpy(GATES[k + "/y"])
        assert util.rel_err(y, ref) < 1e-3
        assert _agree(y.cpu(), ref) >= 0.999


def test_batched_edm_gating_matches_per_image():
    """Per-image gating of a batch (compaction) == the batch-1 reference control flow per image."""
    spec = util.NET_CASES["searched-dense-C2"]
    net = util.make_net(spec).to(DEV)
    edm = util.make_edm().to(DEV)
    x, gt = util.make_input(5, 33, 65, seed=99)
    xd, gtd = x.to(DEV), gt.to(DEV)
    singles = []
    for i in range(5):
        y, ee, _, cv = net.dynamic_inference(xd[i:i + 1], threshold=-1e30, confidence='edm', edm=edm)  # never exit
        singles.append(float(cv))
    srt = sorted(singles)
    thr = 0.5 * (srt[1] + srt[2])            # 2 images exit early, 3 continue
    ref = []
    for i in range(5):
        y, ee, _, cv = net.dynamic_inference(xd[i:i + 1], threshold=thr, confidence='edm', edm=edm)
        ref.append((y.clone(), ee))
    ys, flags, confs = net.dynamic_inference_batch(xd, thr, 'edm', edm)
    assert flags == [r[1] for r in ref] and sum(flags) == 2
    for i in range(5):
        assert util.rel_err(ys[i], ref[i][0]) < 1e-6
        assert float(confs[i]) == pytest.approx(singles[i], rel=1e-5, abs=1e-6)
    cms, flags2, _ = net.dynamic_evaluate(xd, gtd, thr, edm)
    assert flags2 == flags
    for i in range(5):
        want = orc.generate_matrix(gt[i].numpy(), ref[i][0].argmax(1).cpu().numpy())
        assert np.array_equal(cms[i].cpu().numpy(), want)
    # all exit / none exit extremes
    _, f_all, _ = net.dynamic_inference_batch(xd, 1e30, 'edm', edm)
    _, f_none, _ = net.dynamic_inference_batch(xd, -1e30, 'edm', edm)
    assert f_all == [1] * 5 and f_none == [0] * 5


def test_host_pipeline_matches_direct_calls():
    """HostPipeline (pinned host batches, double-buffered H2D, plans bound to the slot buffers) returns exactly what
    direct ADD.dynamic_evaluate / ADD.evaluate calls return, batch after batch (slots are reused across batches)."""
    spec = util.NET_CASES["searched-dense-C2"]
    net = util.make_net(spec).to(DEV)
    edm = util.make_edm().to(DEV)
    batches = [util.make_input(3, 33, 65, seed=300 + i) for i in range(5)]
    _, _, confs = net.dynamic_evaluate(batches[0][0].to(DEV), batches[0][1].to(DEV), -1e30, edm)
    thr = sorted(float(c) for c in confs)[1]
    want = []
    for x, gt in batches:
        cm, flags, _ = net.dynamic_evaluate(x.to(DEV), gt.to(DEV), thr, edm)
        want.append((cm.cpu().clone(), list(flags)))
    pipe = add_b200.HostPipeline(net, edm, thr)
    got = [(cm.clone(), list(flags)) for cm, flags in pipe.evaluate((x.pin_memory(), gt.pin_memory()) for x, gt in batches)]
    assert len(got) == len(want)
    for (cm_g, fl_g), (cm_w, fl_w) in zip(got, want):
        assert fl_g == fl_w
        assert torch.equal(cm_g[0], cm_w)
    # uint8 host labels (1 byte per pixel over PCIe, widened on the device): same matrices
    got8 = [(cm.clone(), list(flags)) for cm, flags in
            pipe.evaluate((x.pin_memory(), gt.to(torch.uint8).pin_memory()) for x, gt in batches)]
    for (cm_g, fl_g), (cm_w, fl_w) in zip(got8, want):
        assert fl_g == fl_w and torch.equal(cm_g[0], cm_w)
    # multi-exit mode (no EDM): per-exit confusion matrices
    pipe2 = add_b200.HostPipeline(net)
    for (cm_g, fl), (x, gt) in zip(pipe2.evaluate((x.pin_memory(), gt.pin_memory()) for x, gt in batches[:2]), batches[:2]):
        assert fl is None
        assert torch.equal(cm_g, net.evaluate(x.to(DEV), gt.to(DEV)).cpu())


def test_dynamic_evaluate_uint8_labels_equal_int64_labels():
    """uint8 labels (the PNG bytes, 255 = ignore) give the same per-image confusion matrices as the reference's int64
    tensors — odd image size (33 x 65 = 2145 bytes per label slab: the byte-granular gather) and an even one."""
    spec = util.NET_CASES["searched-dense-C2"]
    net = util.make_net(spec).to(DEV)
    edm = util.make_edm().to(DEV)
    for hw, seed in (((33, 65), 700), ((48, 80), 701)):
        x, gt = util.make_input(3, *hw, seed=seed)
        _, _, confs = net.dynamic_evaluate(x.to(DEV), gt.to(DEV), -1e30, edm)
        thr = sorted(float(c) for c in confs)[1]
        cm64, fl64, _ = net.dynamic_evaluate(x.to(DEV), gt.to(DEV), thr, edm)
        cm8, fl8, _ = net.dynamic_evaluate(x.to(DEV), gt.to(torch.uint8).to(DEV), thr, edm)
        assert list(fl64) == list(fl8) and 0 < sum(fl64) < 3
        assert torch.equal(cm64, cm8)
        assert int(cm8.sum()) == int((gt != 255).sum())


def test_host_pipeline_uint8_images():
    """uint8 HWC host images (3 bytes per pixel over PCIe, normalised on the device with the reference loader's arithmetic)
    give exactly the matrices of direct calls on the host-normalised fp32 NCHW tensors."""
    spec = util.NET_CASES["searched-dense-C2"]
    net = util.make_net(spec).to(DEV)
    edm = util.make_edm().to(DEV)
    g = torch.Generator().manual_seed(9)
    imgs = [torch.randint(0, 256, (3, 33, 65, 3), generator=g, dtype=torch.uint8) for _ in range(4)]
    gts = [util.make_input(3, 33, 65, seed=500 + i)[1] for i in range(4)]
    xs = [add_b200.normalize_u8_hwc_host(im) for im in imgs]
    _, _, confs = net.dynamic_evaluate(xs[0].to(DEV), gts[0].to(DEV), -1e30, edm)
    thr = sorted(float(c) for c in confs)[1]
    want = [net.dynamic_evaluate(x.to(DEV), gt.to(DEV), thr, edm) for x, gt in zip(xs, gts)]
    want = [(cm.cpu().clone(), list(flags)) for cm, flags, _ in want]
    pipe = add_b200.HostPipeline(net, edm, thr)
    got = [(cm.clone(), list(flags)) for cm, flags in
           pipe.evaluate((im.pin_memory(), gt.to(torch.uint8).pin_memory()) for im, gt in zip(imgs, gts))]
    assert len(got) == len(want)
    for (cm_g, fl_g), (cm_w, fl_w) in zip(got, want):
        assert fl_g == fl_w and torch.equal(cm_g[0], cm_w)


def test_resident_pipeline_matches_direct_calls():
    """ResidentPipeline (device batches cycling over three stable buffers; the trunk of batch i+1 is enqueued before
    the host reads batch i's gate values) yields exactly what blocking dynamic_evaluate calls return, in order,
    for batches with different exit patterns (7 batches over 3 buffers: every buffer is reused)."""
    spec = util.NET_CASES["searched-dense-C2"]
    net = util.make_net(spec).to(DEV)
    edm = util.make_edm().to(DEV)
    batches = [util.make_input(3, 33, 65, seed=400 + i) for i in range(7)]
    _, _, confs = net.dynamic_evaluate(batches[0][0].to(DEV), batches[0][1].to(DEV), -1e30, edm)
    thr = sorted(float(c) for c in confs)[1]
    want = []
    for x, gt in batches:
        cm, flags, _ = net.dynamic_evaluate(x.to(DEV), gt.to(DEV), thr, edm)
        want.append((cm.cpu().clone(), list(flags)))
    assert len({tuple(f) for _, f in want}) > 1, "the batches should not all take the same exits"
    bufs = [(torch.empty_like(batches[0][0], device=DEV), torch.empty_like(batches[0][1], device=DEV)) for _ in range(3)]

    def feed():
        for i, (x, gt) in enumerate(batches):
            bx, bg = bufs[i % 3]
            bx.copy_(x.to(DEV)); bg.copy_(gt.to(DEV))      # stream-ordered after the last compute that read this buffer
            yield bx, bg

    pipe = add_b200.ResidentPipeline(net, edm, thr)
    got = [(cm.cpu().clone(), list(flags)) for cm, flags in pipe.evaluate(feed())]
    assert len(got) == len(want)
    for (cm_g, fl_g), (cm_w, fl_w) in zip(got, want):
        assert fl_g == fl_w
        assert torch.equal(cm_g, cm_w)


@pytest.mark.parametrize("cname", sorted(util.SIBLING_CASES))
@pytest.mark.parametrize("precision,tol,agree_min", [("fp32", 1e-3, 0.999), ("bf16", 1e-1, 0.95)])
def test_sibling_models(cname, precision, tol, agree_min):
    """Baselin_Model / AutoDeepLab drop-ins (ADD's kernels and plans, non-dense wiring) vs the reference goldens."""
    sibs = np.load(util.ROOT / "tests/golden/siblings.npz")
    spec = util.SIBLING_CASES[cname]
    net = util.make_sibling(spec).to(DEV)
    net.set_precision(precision)
    x, _ = util.make_input(1, *spec["size"])
    outs = net(x.to(DEV))
    if spec["cls"] == "AutoDeepLab":
        assert outs[0] is None
        outs = [outs[1]]
    for e, o in enumerate(outs):
        ref = torch.from_numpy(sibs[f"{cname}/forward/{e}"])
        assert util.rel_err(o, ref) < tol
        assert _agree(o.cpu(), ref) >= agree_min
