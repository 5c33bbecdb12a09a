"""GPU: the training step (reference train.py:216-247) — forward in train mode, mean-over-exits
cross entropy, backward through libadd_b200's backward kernels (csrc/backward.cu), SGD-nesterov — against fixtures
produced by the unmodified reference modules' own `.backward()` (tests/golden/train_ops_grad.npz, train_step.npz).
fp32; tolerances: operator gradients 2e-4 max-norm relative (same arithmetic, different summation order); whole-network
first-step gradients 2e-2 on |sum| per tensor and 5e-2 max-norm on the stored tensors (BatchNorm over the 50 samples of the
stride-32 level is ill-conditioned: the reference differs from its own functional restatement by 1e-6 there on
identical inputs and by 1e-1 one step later)."""
import numpy as np
import pytest
import torch

import util
import add_b200
from add_b200 import training as T

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
GRADS = np.load(util.ROOT / "tests/golden/train_ops_grad.npz")
TRAIN = np.load(util.ROOT / "tests/golden/train_step.npz")


def _train_forward(m, x):
    from add_b200.operations import SepConv, DilConv, ReLUConvBN, _FactorizedReduceBase
    if isinstance(m, SepConv):
        return T.sep_conv(m, x)
    if isinstance(m, DilConv):
        return T.dil_conv(m, x)
    if isinstance(m, ReLUConvBN):
        return T.relu_conv_bn(m, x)
    if isinstance(m, _FactorizedReduceBase):
        return T.factorized_reduce(m, x)
    raise TypeError(type(m))


@pytest.mark.parametrize("name", util.TRAIN_OP_CASES)
def test_operator_backward_matches_reference(name):
    m, x, cot = util.make_op_grad_case(name)
    m = m.to(DEV)
    xd = x.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = _train_forward(m, xd)
    assert util.rel_err(y, torch.from_numpy(GRADS[f"{name}/y"])) < 2e-4
    (y * cot.to(DEV)).sum().backward()
    assert util.rel_err(xd.grad, torch.from_numpy(GRADS[f"{name}/dx"])) < 2e-4, "dx"
    for k, p in m.named_parameters():
        ref = torch.from_numpy(GRADS[f"{name}/grad/{k}"])
        assert p.grad is not None, k
        assert util.rel_err(p.grad, ref) < 2e-4 or float((p.grad.cpu() - ref).abs().max()) < 1e-5, k


def test_cross_entropy_matches_torch():
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(2, 20, 17, 23, generator=g) * 3
    gt = torch.randint(0, 19, (2, 17, 23), generator=g)
    gt[torch.rand(2, 17, 23, generator=g) < 0.2] = 255
    cw = torch.rand(19, generator=g) + 0.5
    for weight in (None, cw):
        ref_in = logits[:, :19].clone().requires_grad_(True)
        ref = torch.nn.functional.cross_entropy(ref_in, gt, weight=weight, ignore_index=255)
        ref.backward()
        ld = logits.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        loss = T.cross_entropy(ld, gt.to(DEV), 19, 255, None if weight is None else weight.to(DEV))
        (loss * 0.5).backward()
        assert float(loss) == pytest.approx(float(ref), rel=1e-5)
        assert util.rel_err(ld.grad[:, :19] * 2.0, ref_in.grad) < 1e-5
        assert float(ld.grad[:, 19].abs().max()) == 0.0


def test_bilinear_backward_matches_torch():
    g = torch.Generator().manual_seed(6)
    for (h, w, ho, wo) in [(9, 13, 17, 25), (16, 16, 5, 7), (5, 7, 40, 56), (33, 33, 33, 65), (1, 1, 6, 9)]:
        x = torch.randn(2, 8, h, w, generator=g)
        cot = torch.randn(2, 8, ho, wo, generator=g)
        xr = x.clone().requires_grad_(True)
        (torch.nn.functional.interpolate(xr, (ho, wo), mode="bilinear", align_corners=False) * cot).sum().backward()
        xd = x.to(DEV).contiguous(memory_format=torch.channels_last).requires_grad_(True)
        (T.bilinear(xd, (ho, wo)) * cot.to(DEV)).sum().backward()
        assert util.rel_err(xd.grad, xr.grad) < 1e-5, (h, w, ho, wo)


def test_sgd_nesterov_matches_torch():
    g = torch.Generator().manual_seed(7)
    ps = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in [(7, 3, 3, 3), (40,), (19, 256, 1, 1)]]
    ref = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    dev = [torch.nn.Parameter(p.detach().clone().to(DEV)) for p in ps]
    opt_r = torch.optim.SGD(ref, lr=0.05, momentum=0.9, weight_decay=4e-5, nesterov=True)
    opt_d = T.SGD(dev, lr=0.05, momentum=0.9, weight_decay=4e-5, nesterov=True)
    for step in range(3):
        grads = [torch.randn(p.shape, generator=g) for p in ps]
        opt_r.zero_grad(); opt_d.zero_grad()
        for p, q, gr in zip(ref, dev, grads):
            p.grad = gr.clone()
            q.grad.copy_(gr)
        opt_r.step(); opt_d.step()
        for p, q in zip(ref, dev):
            assert util.rel_err(q.detach(), p.detach()) < 1e-6, step


def test_train_step_matches_reference():
    spec = util.TRAIN_STEP
    net, x, gt = util.make_train_case()
    net = net.to(DEV).train()
    opt = T.SGD(net.parameters(), lr=spec["lr"], momentum=spec["momentum"], weight_decay=spec["weight_decay"], nesterov=spec["nesterov"])
    xd, gtd = x.to(DEV), gt.to(DEV)
    want_abs = {str(k): float(a) for k, a in zip(TRAIN["grad_names"], TRAIN["grad_abs_sum"])}
    # how far the REFERENCE's own gradient of each tensor moves under a 1e-6 relative input perturbation (max-norm): the
    # noise floor of this chaotic random-init network; our deviation must stay within 3x of it (+ 1e-3)
    floor = {str(k): float(a) for k, a in zip(TRAIN["grad_names"], TRAIN["grad_rel_change_pert"])}
    for step in range(spec["steps"]):
        opt.zero_grad()
        loss, outs = T.add_loss(net, xd, gtd)
        assert float(loss) == pytest.approx(float(TRAIN[f"step{step}/loss"]), rel=1e-4 if step == 0 else 5e-3)
        loss.backward()
        if step == 0:
            bad = []
            for k, p in net.named_parameters():
                got = float(p.grad.double().abs().sum())
                if abs(got - want_abs[k]) > (3 * floor[k] + 1e-3) * want_abs[k] + 1e-6:
                    bad.append((k, got, want_abs[k], floor[k]))
                if k in util.TRAIN_FULL_GRADS:
                    err = util.rel_err(p.grad, torch.from_numpy(TRAIN[f"grad/{k}"]))
                    assert err < 3 * floor[k] + 1e-3, (k, err, floor[k])
            assert not bad, bad[:10]
        opt.step()
        sd = net.state_dict()
        tol = 1e-2 if step == 0 else 1e-1      # parameter change = lr x a gradient that is only defined to ~10 % (see floor)
        for k, a in zip(TRAIN[f"step{step}/state_names"], TRAIN[f"step{step}/state_abs_sum"]):
            assert float(sd[str(k)].double().abs().sum()) == pytest.approx(float(a), rel=tol, abs=1e-5), (step, str(k))
    # the step is deterministic: the same two steps again from the same start give bit-identical parameters
    net2, _, _ = util.make_train_case()
    net2 = net2.to(DEV).train()
    opt2 = T.SGD(net2.parameters(), lr=spec["lr"], momentum=spec["momentum"], weight_decay=spec["weight_decay"], nesterov=spec["nesterov"])
    for step in range(spec["steps"]):
        T.train_step(net2, opt2, xd, gtd)
    for (k, a), (_, b) in zip(net.state_dict().items(), net2.state_dict().items()):
        assert torch.equal(a, b), k
    # after training steps the inference plans see the new parameters (eval mode, fused path)
    net.eval()
    out = net(xd)
    assert all(torch.isfinite(o).all() for o in out)


def test_graphed_train_step_equals_eager():
    """The whole iteration captured as one CUDA graph replays to the same parameters as the eager step, bit for bit."""
    spec = util.TRAIN_STEP
    res = []
    for graphed in (False, True):
        net, x, gt = util.make_train_case()
        net = net.to(DEV).train()
        opt = T.SGD(net.parameters(), lr=spec["lr"], momentum=0.9, weight_decay=4e-5, nesterov=True)
        step = T.GraphedTrainStep(net, opt, warmup=1) if graphed else (lambda a, b, lr=None: T.train_step(net, opt, a, b, lr))
        xd, gtd = x[:, :, :65, :65].contiguous().to(DEV), gt[:, :65, :65].contiguous().to(DEV)
        losses = [float(step(xd, gtd, lr=spec["lr"] * (1 - i / 10) ** 0.9)) for i in range(4)]
        res.append((losses, {k: v.clone() for k, v in net.state_dict().items()}))
    assert res[0][0] == res[1][0], (res[0][0], res[1][0])
    for k in res[0][1]:
        assert torch.equal(res[0][1][k], res[1][1][k]), k
