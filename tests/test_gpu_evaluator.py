"""GPU: Evaluator / confusion-matrix histogram — bit-exact against the reference's golden matrices
and the numpy oracle, including edge cases and a full-size property test."""
import numpy as np
import pytest
import torch

import util
import add_b200
from util import orc

pytestmark = pytest.mark.gpu
OPS = np.load(util.ROOT / "tests/golden/ops.npz")
DEV = "cuda:0"


@pytest.mark.parametrize("cname", sorted(util.make_evaluator_cases()))
def test_generate_matrix_bit_exact(cname):
    gt, pred = util.make_evaluator_cases()[cname]
    ev = add_b200.Evaluator(19)
    cm = ev._generate_matrix(gt.to(DEV), pred.to(DEV))
    assert cm.dtype == torch.int64 and tuple(cm.shape) == (19, 19)
    assert np.array_equal(cm.cpu().numpy(), OPS[f"evaluator/{cname}/cm"])
    ev.add_batch(gt.to(DEV), pred.to(DEV))
    assert ev.confusion_matrix.dtype == torch.float32          # Q6: fp32 accumulator kept
    miou, ref = ev.Mean_Intersection_over_Union(), float(OPS[f"evaluator/{cname}/miou"])
    assert (np.isnan(miou) and np.isnan(ref)) or miou == pytest.approx(ref, rel=1e-6)


def test_full_size_properties():
    """8 x 1024 x 2048 (BASELINE config 2 size): matrix total == number of valid pixels, equals the
    sum of per-image matrices (additivity), and equals numpy bincount."""
    g = torch.Generator().manual_seed(77)
    gt = torch.randint(0, 19, (8, 1024, 2048), generator=g)
    gt[torch.rand(8, 1024, 2048, generator=g) < 0.1] = 255
    pred = torch.randint(0, 19, (8, 1024, 2048), generator=g)
    ev = add_b200.Evaluator(19)
    gtd, prd = gt.to(DEV), pred.to(DEV)
    cm = ev._generate_matrix(gtd, prd).cpu().numpy()
    assert cm.sum() == int((gt != 255).sum())
    parts = sum(ev._generate_matrix(gtd[i], prd[i]).cpu().numpy() for i in range(8))
    assert np.array_equal(cm, parts)
    assert np.array_equal(cm, orc.generate_matrix(gt.numpy(), pred.numpy()))
    assert np.array_equal(cm.sum(1), np.bincount(gt[gt != 255].numpy(), minlength=19))


def test_odd_lengths_and_unaligned_tail():
    g = torch.Generator().manual_seed(78)
    for n in (1, 2, 3, 63, 64, 65, 4097, 100001):
        gt = torch.randint(-2, 21, (n,), generator=g)
        pred = torch.randint(0, 19, (n,), generator=g)
        ev = add_b200.Evaluator(19)
        cm = ev._generate_matrix(gt.to(DEV), pred.to(DEV)).cpu().numpy()
        assert np.array_equal(cm, orc.generate_matrix(gt.numpy(), pred.numpy())), n
