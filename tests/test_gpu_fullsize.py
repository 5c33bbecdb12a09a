"""GPU: the BASELINE.json bench configuration (searched-dense C=2, 1024x2048, bf16 tensor-core path, CUDA-graph
replay) checked through size-independent properties — the CPU oracle would need minutes at this size:
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
