"""CPU: the WIRING of the training path (`training.add_forward` / `add_loss`, `ADD.forward` in `.train()`) against the
train-step fixtures of the unmodified reference, with the primitives replaced by plain-PyTorch stand-ins
(tests/sim_training.py).  Same fixture, tolerances and noise floor as tests/test_gpu_training.py, which pins the real
kernels and their backward on the B200."""
import numpy as np
import pytest
import torch

import util
import sim_training
from add_b200 import training as T

TRAIN = np.load(util.ROOT / "tests/golden/train_step.npz")


@pytest.fixture(autouse=True)
def _sim(monkeypatch):
    sim_training.install(monkeypatch)
    yield


def _floor():
    return {str(k): float(a) for k, a in zip(TRAIN["grad_names"], TRAIN["grad_rel_change_pert"])}


def test_first_train_step_loss_and_gradients_match_reference():
    net, x, gt = util.make_train_case()
    net.train()
    loss, outs = T.add_loss(net, x, gt)
    assert len(outs) == 2 and all(tuple(o.shape) == (x.shape[0], 20, *x.shape[2:]) for o in outs)
    assert float(loss) == pytest.approx(float(TRAIN["step0/loss"]), rel=1e-4)
    loss.backward()
    want_abs = {str(k): float(a) for k, a in zip(TRAIN["grad_names"], TRAIN["grad_abs_sum"])}
    floor = _floor()
    bad = []
    for k, p in net.named_parameters():
        got = float(p.grad.double().abs().sum())
        if abs(got - want_abs[k]) > (3 * floor[k] + 1e-3) * want_abs[k] + 1e-6:
            bad.append((k, got, want_abs[k], floor[k]))
        if k in util.TRAIN_FULL_GRADS:
            assert util.rel_err(p.grad, torch.from_numpy(TRAIN[f"grad/{k}"])) < 3 * floor[k] + 1e-3, k
    assert not bad, bad[:10]


def test_model_call_in_train_mode_is_the_training_forward():
    """train.py:227-240 as written: output = model(image) in .train(); torch's criterion and optimiser; .eval() afterwards
    sees the updated parameters (generation bumped)."""
    from add_b200 import runtime as rt
    net, x, gt = util.make_train_case()
    net.train()
    outs = net(x)
    assert all(tuple(o.shape) == (x.shape[0], 19, *x.shape[2:]) and o.requires_grad for o in outs)
    crit = torch.nn.CrossEntropyLoss(ignore_index=255)
    loss = sum(crit(o, gt) for o in outs) / len(outs)
    assert float(loss) == pytest.approx(float(TRAIN["step0/loss"]), rel=1e-4)
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=4e-5, nesterov=True)
    opt.zero_grad()
    loss.backward()
    assert sum(p.grad is not None for p in net.parameters()) > 1000
    g_before = rt.generation()
    opt.step()
    net.eval()
    assert rt.generation() > g_before
    with pytest.raises(NotImplementedError):
        net.train().evaluate(x, gt)
