"""GPU: the BASELINE.json bench configuration (searched-dense C=2, 1024x2048, bf16 tensor-core path, CUDA-graph
replay): (1) pinned to the CPU oracle at full size on BN-calibrated weights (the oracle needs ~1.5 s per image and exit
path here) with stated bf16 tolerances, and (2) checked through size-independent properties:
 * every valid ground-truth pixel is counted exactly once per exit (row sums of the confusion matrix = per-class
   pixel counts of gt; ignored pixels never counted);
 * the fused evaluate path (upsample+argmax+histogram, no full-resolution logits) equals forward -> argmax ->
   Evaluator on the materialised logits, bit for bit;
 * per-image early exit of a batch is batch-composition invariant: an image's exit decision, gate value and
   confusion matrix are the same alone and inside a batch (images never mix in any kernel)."""
import numpy as np
import pytest
import torch

import add_b200

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
H, W = 1024, 2048


@pytest.fixture(scope="module")
def bench_net():
    net = add_b200.build_add("searched-dense", 2, 20, seed=1).to(DEV)
    net.set_precision("bf16")
    net.use_cuda_graph = True
    torch.manual_seed(203)
    edm = add_b200.EDM().eval().to(DEV)
    return net, edm


def _class_counts(gt):
    g = gt.reshape(gt.shape[0], -1)
    return torch.stack([torch.bincount(row[(row >= 0) & (row < 19)], minlength=19) for row in g])


def test_fullsize_every_valid_pixel_counted_once(bench_net):
    net, _ = bench_net
    x, gt = add_b200.synthetic_batch(2, H, W, seed=7)
    cm = net.evaluate(x.to(DEV), gt.to(DEV)).cpu()             # [exits, N, 19, 19]
    want = _class_counts(gt)
    for e in range(cm.shape[0]):
        assert torch.equal(cm[e].sum(2), want)                 # rows = gt classes
        assert int(cm[e].sum()) == int(((gt >= 0) & (gt < 19)).sum())


def test_fullsize_fused_evaluate_equals_forward_argmax_evaluator(bench_net):
    net, _ = bench_net
    x, gt = add_b200.synthetic_batch(1, H, W, seed=8)
    xd, gtd = x.to(DEV), gt.to(DEV)
    cm = net.evaluate(xd, gtd).clone()
    outs = net(xd)
    for e, o in enumerate(outs):
        ev = add_b200.Evaluator(19)
        ev.add_batch(gtd, torch.argmax(o, 1))
        assert torch.equal(ev.confusion_matrix_int64, cm[e].sum(0))


def test_fullsize_early_exit_is_batch_composition_invariant(bench_net):
    net, edm = bench_net
    x, gt = add_b200.synthetic_batch(4, H, W, seed=9)
    xd, gtd = x.to(DEV), gt.to(DEV)
    _, _, confs = net.dynamic_evaluate(xd, gtd, -1e30, edm)
    vals = sorted(float(c) for c in confs)
    thr = 0.5 * (vals[1] + vals[2])
    cm_b, flags_b, confs_b = net.dynamic_evaluate(xd, gtd, thr, edm)
    cm_b = cm_b.clone()
    assert sum(flags_b) == 2
    for i in range(4):
        cm_1, flags_1, confs_1 = net.dynamic_evaluate(xd[i:i + 1], gtd[i:i + 1], thr, edm)
        assert flags_1[0] == flags_b[i]
        assert float(confs_1[0]) == float(confs_b[i])
        assert torch.equal(cm_1[0], cm_b[i])


def test_config1_512x1024_fp32_parity_vs_oracle():
    """BASELINE.json configs[0]: searched-dense, one 3x512x1024 image, fp32, all exits + confusion matrix — the
    CUDA fp32 path against the CPU oracle on the same seeded input and weights.  Gates (north_star): logits within
    1e-3 max-norm relative, argmax agreement >= 99.9 %, confusion matrix bit-exact given identical predictions."""
    from oracle import add_oracle as orc
    net = add_b200.build_add("searched-dense", 2, 20, seed=1)
    sd = orc.randomize_bn_({k: v.clone() for k, v in net.state_dict().items()}, 21)
    net.load_state_dict(sd)
    x, gt = orc.synthetic_batch(1, 512, 1024)
    na, ci, low = add_b200.NETWORKS["searched-dense"][2]
    with torch.no_grad():
        ref = orc.add_forward(sd, orc.Arch(na, ci, low_level_layer=low), x)
    net = net.to(DEV).eval()
    outs = net(x.to(DEV))
    cm = net.evaluate(x.to(DEV), gt.to(DEV)).cpu().numpy()
    for e, (o, r) in enumerate(zip(outs, ref)):
        oc = o.cpu()
        err = float((oc.double() - r.double()).abs().max() / r.double().abs().max())
        agree = float((oc.argmax(1) == r.argmax(1)).float().mean())
        assert err < 1e-3 and agree >= 0.999, (e, err, agree)
        want = orc.generate_matrix(gt.numpy(), oc.argmax(1).numpy())
        assert np.array_equal(cm[e].sum(0), want)


# ---- BASELINE config 2 (1024x2048, bf16 tensor-core path, CUDA graphs) against the CPU oracle ----------------------
# What bf16 can and cannot do here (measured r4d, tools/bf16_parity_probe.py, profiles/r4d_bf16_parity.md):
#  * every activation is stored in bf16 (2^-9 relative rounding) at ~60 / ~120 sequential roundings along the deepest
#    path to the first / last exit; a random-init network has no confident predictions, so its two best logits are
#    often closer than that noise and the argmax of such pixels flips.  North_star's 99.9 % argmax agreement is met by
#    the fp32 CUDA path (tests above / test_gpu_net.py: rel 2e-5, argmax 99.99-100 %); for bf16 the tolerance is STATED
#    per exit and weight set, and agreement is stated overall and on the DECISIVE pixels (fp32 top-2 logit gap above the
#    stated tolerance x the largest logit), where it must be >= 99.9 %;
#  * the error is inherent to bf16 storage, not to these kernels: stock PyTorch bf16 (cuDNN, `model.bfloat16()`) on
#    the same weights and input is measured in the same test and our rms error must not exceed it (measured 0.7-0.9x:
#    BN is folded before rounding and SepConv / node sums round less often).
# Weight sets: "randomized" = tests/util's BN-randomised statistics (activations grow with depth, no cancellation);
# "calibrated" = one training-mode forward of the oracle writes every running statistic (activations O(1),
# every BN subtracts a mean, so rounding noise is amplified relative to what is left).
BF16_TOL = {            # (forward first exit, forward last exit, dynamic_inference early exit, dynamic_inference last exit):
                        # max-norm relative tolerance, overall argmax-agreement floor
    "randomized": dict(rel=(4e-2, 2e-1, 5e-2, 2e-1), agree=(0.97, 0.87, 0.97, 0.85)),
    "calibrated": dict(rel=(2e-1, 5e-1, 2.5e-1, 5e-1), agree=(0.88, 0.78, 0.84, 0.72)),
}


def _parity(o, r, tol):
    o, r = o.double().cpu(), r.double()
    scale = r.abs().max()
    agree = o.argmax(1) == r.argmax(1)
    top2 = r.topk(2, 1).values
    dec = (top2[:, 0] - top2[:, 1]) / scale > tol
    return dict(rel=float((o - r).abs().max() / scale), rms=float((o - r).pow(2).mean().sqrt() / scale),
                agree=float(agree.float().mean()), agree_decisive=float(agree[dec].float().mean()) if dec.any() else 1.0,
                frac_decisive=float(dec.float().mean()))


def _torch_bf16_forward(orc, sd, arch, x):
    """Stock PyTorch bf16 on the GPU (cuDNN / ATen): the oracle's functional graph with bf16 weights + activations."""
    def cast(v):
        v = v.to(DEV)
        if v.is_floating_point():
            v = v.to(torch.bfloat16)
            if v.dim() == 4:
                v = v.contiguous(memory_format=torch.channels_last)
        return v
    with torch.no_grad():
        return [o.float() for o in orc.add_forward({k: cast(v) for k, v in sd.items()}, arch, cast(x))]


@pytest.fixture(scope="module", params=["randomized", "calibrated"])
def weights(request):
    """searched-dense weights (randomised or calibrated BN statistics), the bench path's precision / graph settings, a
    seeded EDM."""
    from oracle import add_oracle as orc
    na, ci, low = add_b200.NETWORKS["searched-dense"][2]
    arch = orc.Arch(na, ci, low_level_layer=low)
    net = add_b200.build_add("searched-dense", 2, 20, seed=1)
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    sd = orc.randomize_bn_(sd, 21) if request.param == "randomized" else orc.calibrate_bn_(sd, arch)
    net.load_state_dict(sd)
    net = net.to(DEV).eval()
    net.set_precision("bf16")
    net.use_cuda_graph = True
    torch.manual_seed(203)
    edm = add_b200.EDM().eval()
    edm_sd = {k: v.detach().clone() for k, v in edm.state_dict().items()}
    return request.param, orc, arch, sd, net, edm.to(DEV), edm_sd


def test_config2_bf16_forward_vs_oracle(weights):
    """ADD.forward, all exits, one 1024x2048 image: the bf16 CUDA-graph path against the fp32 CPU oracle."""
    wname, orc, arch, sd, net, _, _ = weights
    tol = BF16_TOL[wname]
    x, gt = orc.synthetic_batch(1, H, W, seed=4321)
    with torch.no_grad():
        ref = orc.add_forward(sd, arch, x)
    outs = net(x.to(DEV))
    stock = _torch_bf16_forward(orc, sd, arch, x)
    cm = net.evaluate(x.to(DEV), gt.to(DEV)).cpu().numpy()
    for e, (o, r, t) in enumerate(zip(outs, ref, stock)):
        m, ms = _parity(o, r, tol["rel"][e]), _parity(t, r, tol["rel"][e])
        print(f"config2 bf16 forward [{wname}] exit {e}: ours {m} | stock PyTorch bf16 rms {ms['rms']:.3e} agree {ms['agree']:.4f}")
        assert m["rel"] < tol["rel"][e], (e, m)
        assert m["agree"] >= tol["agree"][e], (e, m)
        assert m["agree_decisive"] >= 0.999, (e, m)
        assert m["rms"] <= 1.05 * ms["rms"], (e, m, ms)            # no avoidable precision loss against cuDNN bf16
        # integer contract: the fused head's confusion matrix is bit-exact given ITS OWN predictions
        want = orc.generate_matrix(gt.numpy(), o.cpu().argmax(1).numpy())
        assert np.array_equal(cm[e].sum(0), want)


@pytest.mark.parametrize("label", ["exit", "noexit"])
def test_config2_bf16_dynamic_inference_vs_oracle(weights, label):
    """ADD.dynamic_inference (EDM gate, reference exit semantics: the early exit runs ASPP on the x4 up-sampled map,
    SURVEY Q3) at 1024x2048: decision, gate value and logits of the bf16 path against the oracle; the fused
    dynamic_evaluate confusion matrix is bit-exact given the path's own predictions."""
    wname, orc, arch, sd, net, edm, edm_sd = weights
    tol = BF16_TOL[wname]
    x, gt = orc.synthetic_batch(1, H, W, seed=4322)
    with torch.no_grad():
        _, _, c0 = orc.add_dynamic_inference(sd, arch, x, -1e30, 'edm', edm_sd)
        thr = float(c0) + (1.0 if label == "exit" else -1.0) * max(1.0, 0.1 * abs(float(c0)))
        y_ref, ee_ref, cv_ref = orc.add_dynamic_inference(sd, arch, x, thr, 'edm', edm_sd)
    y, ee, _, cv = net.dynamic_inference(x.to(DEV), threshold=thr, confidence='edm', edm=edm)
    assert ee == ee_ref == (1 if label == "exit" else 0)
    assert float(cv) == pytest.approx(float(cv_ref), rel=2e-2, abs=2e-3)         # stated bf16 margin of the gate value
    e = 2 if label == "exit" else 3
    m = _parity(y, y_ref, tol["rel"][e])
    print(f"config2 bf16 dynamic_inference [{wname}] {label}: {m} gate {float(cv):.6f} vs {float(cv_ref):.6f}")
    assert m["rel"] < tol["rel"][e] and m["agree"] >= tol["agree"][e] and m["agree_decisive"] >= 0.999, m
    cmd, flags, _ = net.dynamic_evaluate(x.to(DEV), gt.to(DEV), thr, edm)
    assert flags == [ee]
    assert np.array_equal(cmd[0].cpu().numpy(), orc.generate_matrix(gt.numpy(), y.cpu().argmax(1).numpy()))


def test_edm_gate_decisions_bf16_vs_fp32(weights):
    """EDM gate decisions of the bf16 path against the fp32 oracle over a batch (identical except within a
    stated margin of the threshold).  Margin: |gate_bf16 - gate_fp32| <= 2e-2 * max(|gate|, 0.1)."""
    wname, orc, arch, sd, net, edm, edm_sd = weights
    n, h, w = 8, 257, 513
    x, _ = orc.synthetic_batch(n, h, w, seed=99)
    ref = []
    with torch.no_grad():
        for i in range(n):
            _, feat = orc.add_get_feature(sd, arch, x[i:i + 1])
            ref.append(float(orc.edm_forward(edm_sd, feat.clone())))
    _, _, confs = net.dynamic_inference_batch(x.to(DEV), -1e30, 'edm', edm)
    got = [float(c) for c in confs]
    margin = [2e-2 * max(abs(r), 0.1) for r in ref]
    print(f"EDM gate values [{wname}] bf16 {got} fp32 {ref}")
    for g, r, mg in zip(got, ref, margin):
        assert abs(g - r) <= mg, (got, ref)
    srt = sorted(ref)
    for thr in [0.5 * (srt[i] + srt[i + 1]) for i in range(n - 1)]:
        _, flags, _ = net.dynamic_inference_batch(x.to(DEV), thr, 'edm', edm)
        for f, g, r, mg in zip(flags, got, ref, margin):
            if abs(r - thr) > mg:                       # outside the margin the decision must be the reference's
                assert f == (0 if r > thr else 1), (thr, got, ref)
