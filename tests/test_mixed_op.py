"""MixedOp (cell_level_search.py:10-29): sum_k w_k * op_k(x) over the eight primitives.
Fixtures: the unmodified reference module (tests/golden/mixed_op.npz) in eval mode, in training mode (batch statistics)
and on the argmax path.  CPU: oracle vs fixtures, state_dict keys.  GPU: the drop-in vs fixtures (fp32 1e-4, bf16 3e-2)."""
import numpy as np
import pytest
import torch

import util
import add_b200
from util import orc

GOLD = np.load(util.ROOT / "tests/golden/mixed_op.npz")
CASES = sorted(util.MIXED_OP_CASES)


def _sd(m):
    return {f"m.{k}": v.detach().clone() for k, v in m.state_dict().items()}


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_fixture(name):
    m, x, w = util.make_mixed_op_case(name)
    assert np.array_equal(GOLD[name + "/w"], w.numpy())
    assert util.rel_err(orc.mixed_op(_sd(m), "m", x, w), torch.from_numpy(GOLD[name + "/eval/y"])) < 2e-5
    assert util.rel_err(orc.mixed_op(_sd(m), "m", x, w, training=False), torch.from_numpy(GOLD[name + "/eval/y_argmax"])) < 2e-5
    sd = _sd(m)
    with orc.bn_training(0.1):
        y = orc.mixed_op(sd, "m", x, w)
    assert util.rel_err(y, torch.from_numpy(GOLD[name + "/train/y"])) < 2e-5
    for k in [k for k in GOLD.files if k.startswith(name + "/train/sd/")]:
        assert util.rel_err(sd["m." + k.split("/sd/")[1]], torch.from_numpy(GOLD[k])) < 2e-5


def test_primitive_order_and_keys():
    assert add_b200.PRIMITIVES == ['none', 'max_pool_3x3', 'avg_pool_3x3', 'skip_connect', 'sep_conv_3x3', 'sep_conv_5x5',
                                   'dil_conv_3x3', 'dil_conv_5x5']            # modeling/genotypes.py:5-14
    m = add_b200.MixedOp(16, 1, torch.nn.BatchNorm2d)
    keys = list(m.state_dict())
    assert len(keys) == 34 and "_ops.1.1.running_mean" in keys and "_ops.4.op.2.weight" in keys
    assert not any(k.endswith(".bias") or k.endswith("op.3.weight") for k in keys)      # affine=False everywhere


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 3e-2)], ids=["fp32", "bf16"])
def test_gpu_mixed_op(name, dtype, tol):
    m, x, w = util.make_mixed_op_case(name)
    m = m.to("cuda:0")
    xd = x.to("cuda:0").to(dtype)
    m.eval()
    assert util.rel_err(m(xd, w).float(), torch.from_numpy(GOLD[name + "/eval/y"])) < tol
    assert util.rel_err(m(xd, w, training=False).float(), torch.from_numpy(GOLD[name + "/eval/y_argmax"])) < tol
    m.train()
    y = m(xd, w.to("cuda:0"))
    assert util.rel_err(y.float(), torch.from_numpy(GOLD[name + "/train/y"])) < tol
    for k in [k for k in GOLD.files if k.startswith(name + "/train/sd/")]:
        assert util.rel_err(m.state_dict()[k.split("/sd/")[1]], torch.from_numpy(GOLD[k])) < tol
