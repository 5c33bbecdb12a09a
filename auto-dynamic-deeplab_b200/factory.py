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


CITYSCAPES_MEAN = (0.29866842, 0.30135223, 0.30561872)     # dataloaders/datasets/cityscapes.py:53
CITYSCAPES_STD = (0.23925215, 0.23859318, 0.2385942)       # dataloaders/datasets/cityscapes.py:54


def normalize_u8_hwc_host(img_u8: torch.Tensor, mean=CITYSCAPES_MEAN, std=CITYSCAPES_STD) -> torch.Tensor:
    """Host-side data synthesis helper: what the reference's loader makes of a uint8 [N,H,W,3] image batch
    (custom_transforms.py:17-24, :39) — /255 in float32, -mean and /std through float64 — as fp32 NCHW."""
    v = img_u8.to(torch.float32) / 255.0
    v = (v.double() - torch.tensor(mean, dtype=torch.float64)).float()
    v = (v.double() / torch.tensor(std, dtype=torch.float64)).float()
    return v.permute(0, 3, 1, 2).contiguous()


def synthetic_batch_u8(n: int, h: int, w: int, seed: int = 1234, num_class: int = 19, pin: bool = False):
    """Synthetic Cityscapes-shaped batch as the PNG decoder delivers it: uint8 HWC images whose normalised values are
    ≈ N(0,1) (clipped to the uint8 range), uint8 labels (255 = ignore).  Returns (img_u8 [N,H,W,3], x fp32 NCHW =
    the reference loader's normalisation of img_u8, gt int64 [N,H,W])."""
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(n, h, w, 3, generator=g)
    img = ((z * torch.tensor(CITYSCAPES_STD) + torch.tensor(CITYSCAPES_MEAN)) * 255.0).round().clamp_(0, 255).to(torch.uint8)
    gt = torch.randint(0, num_class, (n, h, w), generator=g, dtype=torch.int64)
    gt[torch.rand(n, h, w, generator=g) < 0.1] = 255
    x = normalize_u8_hwc_host(img)
    if pin:
        img, x, gt = img.pin_memory(), x.pin_memory(), gt.pin_memory()
    return img, x, gt
