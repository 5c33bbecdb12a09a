"""Training-mode forward of the conv operators: ReLU -> conv(s) with the raw weights -> BatchNorm on
BATCH statistics, running statistics updated.  Fixtures: the unmodified reference modules in `.train()` on one device
(tests/golden/train_ops.npz).  CPU: the oracle under `bn_training` against the fixtures.  GPU: the drop-in modules in
`.train()` (raw-weight conv kernels + add_bn_stats_fwd / add_bn_finalize / add_bn_apply_fwd) against the same fixtures,
fp32 1e-4 max-norm relative (outputs and running statistics), bf16 activations 3e-2."""
import numpy as np
import pytest
import torch

import util
from util import orc

GOLD = np.load(util.ROOT / "tests/golden/train_ops.npz")


def _oracle(name, m, x):
    spec = util.OP_CASES[name]
    sd = {f"m.{k}": v.detach().clone() for k, v in m.state_dict().items()}
    kind, args = spec["kind"], spec["args"]
    with orc.bn_training(0.1):
        if kind == "OPS":
            y = orc.apply_primitive(sd, "m", args[0], x, 1)
        elif kind == "ReLUConvBN":
            y = orc.relu_conv_bn(sd, "m", x, args[3], args[4])
        else:
            y = orc.factorized_reduce(sd, "m", x, 2 if kind == "FactorizedReduce" else 4)
    return y, sd


@pytest.mark.parametrize("name", util.TRAIN_OP_CASES)
def test_oracle_training_forward_matches_reference(name):
    m, x = util.make_op_case(name)
    y, sd = _oracle(name, m, x)
    assert util.rel_err(y, torch.from_numpy(GOLD[name + "/y"])) < 2e-5
    keys = [k for k in GOLD.files if k.startswith(name + "/sd/")]
    assert keys
    for k in keys:
        assert util.rel_err(sd["m." + k.split("/sd/")[1]], torch.from_numpy(GOLD[k])) < 2e-5
    # and the oracle's eval mode is untouched by the context manager
    assert not orc._BN_TRAIN["on"]


@pytest.mark.gpu
@pytest.mark.parametrize("name", util.TRAIN_OP_CASES)
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-4), (torch.bfloat16, 3e-2)], ids=["fp32", "bf16"])
def test_gpu_training_forward_matches_reference(name, dtype, tol):
    m, x = util.make_op_case(name)
    m = m.to("cuda:0").train()
    y = m(x.to("cuda:0").to(dtype))
    assert y.dtype == dtype
    assert util.rel_err(y.float(), torch.from_numpy(GOLD[name + "/y"])) < tol
    for k in [k for k in GOLD.files if k.startswith(name + "/sd/")]:
        got = m.state_dict()[k.split("/sd/")[1]]
        assert util.rel_err(got, torch.from_numpy(GOLD[k])) < tol
    for k, v in m.state_dict().items():
        if k.endswith("num_batches_tracked"):
            assert int(v) == 1
    # back in eval mode the fused (BN-folded) path now uses the updated running statistics
    m.eval()
    y_eval = m(x.to("cuda:0").to(dtype))
    assert torch.isfinite(y_eval.float()).all()
