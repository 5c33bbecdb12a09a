"""Supernet cell and search-time ASPP head (reference modeling/cell_level_search.py:32-155,
operations.py:122-158, alpha softmax model_net_search.py:294-310) against fixtures produced by the unmodified reference
(tests/golden/search_cell.npz): eval forward, training forward, and the backward — gradients of the inputs, of the raw
alphas (through the device softmax) and of every weight.  CPU: the oracle; GPU: the drop-ins (fp32, 2e-4)."""
import numpy as np
import pytest
import torch

import util
import add_b200
from util import orc

G = np.load(util.ROOT / "tests/golden/search_cell.npz")


def test_oracle_search_cell_and_aspp_match_reference():
    c = util.SEARCH_CELL
    m, s0, s1_same, s1_up, alphas, cots = util.make_search_cell_case()
    sd = {f"m.{k}": v.detach().clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        outs = orc.search_cell_forward(sd, "m", c["B"], s0, None, s1_same, s1_up, torch.softmax(alphas, -1))
    for i, o in enumerate(outs):
        assert util.rel_err(o, torch.from_numpy(G[f"cell/eval/out{i}"])) < 2e-5
    with torch.no_grad(), orc.bn_training(0.1):
        outs = orc.search_cell_forward(sd, "m", c["B"], s0, None, s1_same, s1_up, torch.softmax(alphas, -1))
    for i, o in enumerate(outs):
        assert util.rel_err(o, torch.from_numpy(G[f"cell/train/out{i}"])) < 2e-5
    a = util.SEARCH_ASPP
    m, x, cot = util.make_search_aspp_case()
    sd = {f"m.{k}": v.detach().clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        assert util.rel_err(orc.search_aspp(sd, "m", x, a["pad"], a["dil"]), torch.from_numpy(G["aspp/eval/y"])) < 2e-5
        with orc.bn_training(0.1):
            assert util.rel_err(orc.search_aspp(sd, "m", x, a["pad"], a["dil"]), torch.from_numpy(G["aspp/train/y"])) < 2e-5


def test_search_cell_state_dict_keys():
    m, *_ = util.make_search_cell_case()
    keys = list(m.state_dict())
    assert "preprocess_same.op.1.weight" in keys and "preprocess_up.op.2.running_mean" in keys and "pre_preprocess.op.1.weight" in keys
    assert "_ops.0._ops.4.op.1.weight" in keys and "_ops.4._ops.1.1.running_var" in keys
    assert not any(k.startswith("preprocess_down") for k in keys)


@pytest.mark.gpu
def test_gpu_search_cell_forward_backward():
    dev = "cuda:0"
    m, s0, s1_same, s1_up, alphas, cots = util.make_search_cell_case()
    m = m.to(dev)
    m.eval()
    with torch.no_grad():
        outs = m(s0.to(dev), None, s1_same.to(dev), s1_up.to(dev), add_b200.softmax_rows(alphas.to(dev)))
    for i, o in enumerate(outs):
        assert util.rel_err(o, torch.from_numpy(G[f"cell/eval/out{i}"])) < 2e-4, f"eval out{i}"
    m.train()
    ins = [t.to(dev).contiguous(memory_format=torch.channels_last).requires_grad_(True) for t in (s0, s1_same, s1_up)]
    al = alphas.to(dev).requires_grad_(True)
    outs = m(ins[0], None, ins[1], ins[2], add_b200.softmax_rows(al))
    for i, o in enumerate(outs):
        assert util.rel_err(o, torch.from_numpy(G[f"cell/train/out{i}"])) < 2e-4, f"train out{i}"
    sum((o * ct.to(dev)).sum() for o, ct in zip(outs, cots)).backward()
    for nme, t in zip(("s0", "s1_same", "s1_up"), ins):
        assert util.rel_err(t.grad, torch.from_numpy(G[f"cell/train/d_{nme}"])) < 5e-4, nme
    assert util.rel_err(al.grad, torch.from_numpy(G["cell/train/d_alphas"])) < 5e-4
    for k, p in m.named_parameters():
        ref = torch.from_numpy(G[f"cell/train/grad/{k}"])
        assert p.grad is not None, k
        assert util.rel_err(p.grad, ref) < 1e-3 or float((p.grad.cpu() - ref).abs().max()) < 1e-5, k
    for k in [k for k in G.files if k.startswith("cell/train/sd/")]:
        assert util.rel_err(m.state_dict()[k.split("/sd/")[1]], torch.from_numpy(G[k])) < 2e-4, k


@pytest.mark.gpu
def test_gpu_search_aspp_forward_backward():
    dev = "cuda:0"
    m, x, cot = util.make_search_aspp_case()
    m = m.to(dev)
    m.eval()
    assert util.rel_err(m(x.to(dev)), torch.from_numpy(G["aspp/eval/y"])) < 2e-4
    m.train()
    xd = x.to(dev).contiguous(memory_format=torch.channels_last).requires_grad_(True)
    y = m(xd)
    assert util.rel_err(y, torch.from_numpy(G["aspp/train/y"])) < 2e-4
    (y * cot.to(dev)).sum().backward()
    assert util.rel_err(xd.grad, torch.from_numpy(G["aspp/train/dx"])) < 5e-4
    for k, p in m.named_parameters():
        ref = torch.from_numpy(G[f"aspp/train/grad/{k}"])
        assert util.rel_err(p.grad, ref) < 1e-3 or float((p.grad.cpu() - ref).abs().max()) < 1e-5, k


@pytest.mark.gpu
def test_gpu_softmax_rows_matches_torch():
    g = torch.Generator().manual_seed(3)
    x = (torch.randn(20, 8, generator=g) * 2).to("cuda:0").requires_grad_(True)
    y = add_b200.softmax_rows(x)
    cot = torch.randn(20, 8, generator=g).to("cuda:0")
    (y * cot).sum().backward()
    xr = x.detach().clone().requires_grad_(True)
    yr = torch.softmax(xr, -1)
    (yr * cot).sum().backward()
    assert util.rel_err(y, yr) < 1e-6 and util.rel_err(x.grad, xr.grad) < 1e-5
