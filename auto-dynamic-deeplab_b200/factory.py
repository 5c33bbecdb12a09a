"""Convenience constructors mirroring the reference drivers' hard-coded configs (eval.py:42-92)."""
from __future__ import annotations

from types import SimpleNamespace

import torch

from .ADD import ADD
from .genotypes import AUTODEEPLAB_CELL, NETWORKS


def Args(F: int = 20, B: int = 5, sync_bn: bool = False) -> SimpleNamespace:
    """The three fields of the reference's argparse namespace that ADD reads (ADD.py:128-130)."""
    return SimpleNamespace(F=F, B=B, sync_bn=sync_bn)


def build_add(network: str = 'searched-dense', C: int = 2, F: int = 20, B: int = 5, num_classes: int = 19,
              seed: int | None = 1) -> ADD:
    """`eval.py --network <network> --C <C> --F <F>` model, randomly initialised on the CPU under
    `seed` (eval.py:276,302 default seed 1), in eval mode."""
    na, ci, low = NETWORKS[network][C]
    if seed is not None:
        torch.manual_seed(seed)
    model = ADD(na, ci, AUTODEEPLAB_CELL.copy(), num_classes, Args(F, B), low)
    return model.eval()


def synthetic_batch(n: int, h: int, w: int, seed: int = 1234, num_class: int = 19, pin: bool = False):
    """SURVEY §8d synthetic Cityscapes-shaped batch on the HOST: image ~ N(0,1) fp32 NCHW (normalised
    Cityscapes images are ≈N(0,1), cityscapes.py:53-54); labels randint(0,19) int64 with 10 % = 255."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, h, w, generator=g)
    gt = torch.randint(0, num_class, (n, h, w), generator=g, dtype=torch.int64)
    gt[torch.rand(n, h, w, generator=g) < 0.1] = 255
    if pin:
        x, gt = x.pin_memory(), gt.pin_memory()
    return x, gt
