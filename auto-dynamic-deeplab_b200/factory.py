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
