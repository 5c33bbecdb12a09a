"""Loader / dump edges either side of the network — device versions of the reference's host code:

* `encode_segmap`            dataloaders/datasets/cityscapes.py:85-91   Cityscapes label ids -> train ids (255 = ignore)
* `full_image_eval_preprocess`  dataloaders/custom_transforms.py:322-347  ToTensor + Normalize + pad to (1025, 2049)
* `decode_segmap`            dataloaders/utils.py:14-51                 class map -> colour image
* `load_checkpoint`          eval.py:126-140                            {'epoch','state_dict','optimizer','best_pred'}, `module.` strip

The PNG bytes (uint8 HWC image, uint8 label-id map) cross PCIe as they are; the remap / normalise / pad run on the
device in libadd_b200 kernels (csrc/io_edges.cu) with the reference's arithmetic, bit-identical.  No CPU fallback."""
from __future__ import annotations

import ctypes
from collections import OrderedDict
from typing import Optional, Tuple

import numpy as np
import torch

from . import runtime as rt
from ._lib import lib, check

# dataloaders/datasets/cityscapes.py:44-45,52
VOID_CLASSES = [0, 1, 2, 3, 4, 5, 6, 9, 10, 14, 15, 16, 18, 29, 30, -1]
VALID_CLASSES = [7, 8, 11, 12, 13, 17, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 31, 32, 33]
IGNORE_INDEX = 255
CITYSCAPES_MEAN = (0.29866842, 0.30135223, 0.30561872)     # cityscapes.py:53
CITYSCAPES_STD = (0.23925215, 0.23859318, 0.2385942)       # cityscapes.py:54
EVAL_CROP = (1025, 2049)                                   # cityscapes.py:109-119


def cityscapes_label_lut() -> np.ndarray:
    """The 256-entry table equivalent to `encode_segmap` on a uint8 map: the reference rewrites the mask in place,
    first every void id to 255, then every valid id to its train id in ascending order — no rule ever matches a value
    written by an earlier rule (train ids are smaller than the ids still to be processed, and 255 is never a source),
    so the cascade is a plain per-value table.  Ids in neither list (34..254) pass through unchanged."""
    lut = np.arange(256, dtype=np.uint8)
    for v in VOID_CLASSES:
        if 0 <= v <= 255:
            lut[v] = IGNORE_INDEX
    for t, v in enumerate(VALID_CLASSES):
        lut[v] = t
    return lut


def get_cityscapes_labels() -> np.ndarray:
    """dataloaders/utils.py:75-95 (the 19 Cityscapes train-id colours)."""
    return np.array([[128, 64, 128], [244, 35, 232], [70, 70, 70], [102, 102, 156], [190, 153, 153], [153, 153, 153],
                     [250, 170, 30], [220, 220, 0], [107, 142, 35], [152, 251, 152], [0, 130, 180], [220, 20, 60],
                     [255, 0, 0], [0, 0, 142], [0, 0, 70], [0, 60, 100], [0, 80, 100], [0, 0, 230], [119, 11, 32]])


_LUT_CACHE: dict = {}


def _device_table(name: str, make, device) -> torch.Tensor:
    key = (name, str(device))
    if key not in _LUT_CACHE:
        _LUT_CACHE[key] = torch.from_numpy(np.ascontiguousarray(make())).to(device)
    return _LUT_CACHE[key]


def _stream(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def encode_segmap(mask: torch.Tensor, pad_to: Optional[Tuple[int, int]] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Label ids -> train ids on the device.  mask: uint8 CUDA tensor [H,W] or [N,H,W] (the gtFine_labelIds PNG bytes).
    pad_to=(Hp,Wp): also pad bottom / right with 255 like `full_image_eval_preprocess`.  Returns uint8 (the dtype the
    fused head reads, `add_upsample_argmax_u8_fwd`); `.long()` gives the reference's LongTensor."""
    rt.require_cuda(mask, "mask")
    if mask.dtype != torch.uint8:
        raise ValueError(f"encode_segmap takes the uint8 label-id map, got {mask.dtype}")
    m = mask if mask.dim() == 3 else mask.unsqueeze(0)
    m = m.contiguous()
    n, h, w = m.shape
    Hp, Wp = (max(pad_to[0], h), max(pad_to[1], w)) if pad_to is not None else (h, w)
    if out is None:
        out = torch.empty((n, Hp, Wp), dtype=torch.uint8, device=m.device)
    assert tuple(out.shape) == (n, Hp, Wp) and out.dtype == torch.uint8 and out.is_contiguous()
    lut = _device_table("label_lut", cityscapes_label_lut, m.device)
    check(lib.add_encode_pad_labels_u8(m.data_ptr(), out.data_ptr(), n, h, w, Hp, Wp, lut.data_ptr(), IGNORE_INDEX,
                                       _stream(m.device)), "encode_pad_labels_u8")
    return out if mask.dim() == 3 else out[0]


def pad_labels(mask: torch.Tensor, pad_to: Tuple[int, int], fill: int = IGNORE_INDEX) -> torch.Tensor:
    """ConstantPad2d((0, pad_lr, 0, pad_tb), 255) of already encoded uint8 labels (custom_transforms.py:344)."""
    rt.require_cuda(mask, "mask")
    m = (mask if mask.dim() == 3 else mask.unsqueeze(0)).contiguous()
    n, h, w = m.shape
    Hp, Wp = max(pad_to[0], h), max(pad_to[1], w)
    out = torch.empty((n, Hp, Wp), dtype=torch.uint8, device=m.device)
    check(lib.add_encode_pad_labels_u8(m.data_ptr(), out.data_ptr(), n, h, w, Hp, Wp, None, int(fill), _stream(m.device)),
          "encode_pad_labels_u8")
    return out if mask.dim() == 3 else out[0]


class full_image_eval_preprocess(object):
    """custom_transforms.py:322-347 on the device: sample = {'image': uint8 [H,W,3] or [N,H,W,3] CUDA tensor (the decoded
    PNG), 'label': uint8 [H,W] / [N,H,W] train-id map} -> {'image': fp32 [3,Hp,Wp] / [N,3,Hp,Wp] normalised and zero
    padded to crop_size, 'label': uint8 padded with 255}.  One kernel each; arithmetic bit-identical to
    ToTensor + Normalize (+ ZeroPad2d / ConstantPad2d)."""

    def __init__(self, crop_size=EVAL_CROP, mean=CITYSCAPES_MEAN, std=CITYSCAPES_STD):
        self.crop_size, self.mean, self.std = tuple(crop_size), tuple(mean), tuple(std)

    def __call__(self, sample):
        image, mask = sample['image'], sample['label']
        rt.require_cuda(image, "image")
        if image.dtype != torch.uint8 or image.shape[-1] != 3:
            raise ValueError(f"image must be uint8 [..., H, W, 3], got {image.dtype} {tuple(image.shape)}")
        img = (image if image.dim() == 4 else image.unsqueeze(0)).contiguous()
        n, h, w, _ = img.shape
        Hp, Wp = max(self.crop_size[0], h), max(self.crop_size[1], w)
        out = torch.empty((n, 3, Hp, Wp), dtype=torch.float32, device=img.device)
        check(lib.add_normalize_pad_u8_hwc_to_nchw(img.data_ptr(), out.data_ptr(), n, h, w, Hp, Wp, *self.mean, *self.std,
                                                   _stream(img.device)), "normalize_pad_u8")
        lab = pad_labels(mask, (Hp, Wp))
        return {'image': out if image.dim() == 4 else out[0], 'label': lab}


def decode_segmap(label_mask: torch.Tensor, dataset: str = 'cityscapes', as_uint8: bool = False):
    """dataloaders/utils.py:14-51 for dataset in {'cityscapes', 'kd'}: class map [H,W] (int64 argmax output or uint8, CUDA) ->
    colour image.  Returns what the reference returns — a float64 numpy array [H,W,3] with values colour/255 — or, with
    as_uint8, the uint8 [H,W,3] CUDA tensor the kernel wrote (for PNG dumps).  Like the reference, a label outside
    [0, 19) keeps its own value in all three channels (e.g. 255 -> white)."""
    if dataset not in ('cityscapes', 'kd'):
        raise NotImplementedError
    rt.require_cuda(label_mask, "label_mask")
    if label_mask.dtype not in (torch.int64, torch.uint8):
        raise ValueError(f"label_mask must be int64 or uint8, got {label_mask.dtype}")
    m = label_mask.contiguous()

    def table():
        lut = np.repeat(np.arange(256, dtype=np.uint8)[:, None], 3, axis=1)
        lut[:19] = get_cityscapes_labels().astype(np.uint8)
        return lut.reshape(-1)
    lut = _device_table("colour_lut", table, m.device)
    rgb = torch.empty(tuple(m.shape) + (3,), dtype=torch.uint8, device=m.device)
    check(lib.add_decode_segmap(m.data_ptr(), 1 if m.dtype == torch.int64 else 0, rgb.data_ptr(), m.numel(), lut.data_ptr(),
                                _stream(m.device)), "decode_segmap")
    if as_uint8:
        return rgb
    return rgb.cpu().numpy().astype(np.float64) / 255.0


def load_checkpoint(model: torch.nn.Module, checkpoint, clean_module: Optional[bool] = None, strict: bool = True):
    """eval.py:126-140 / train.py:186-207: load a reference checkpoint `{'epoch', 'state_dict', 'optimizer', 'best_pred'}`
    (a path or the dict itself) into a drop-in model.  clean_module=True strips the 7-character 'module.' prefix that
    nn.DataParallel / DistributedDataParallel leave on every key (eval.py:133-136); None = strip iff every key has it.
    A bare state_dict is accepted too.  Returns (epoch, best_pred) (None when absent)."""
    if isinstance(checkpoint, (str, bytes)) or hasattr(checkpoint, "__fspath__"):
        import os
        if not os.path.isfile(checkpoint):
            raise RuntimeError("=> no checkpoint found at '{}'".format(checkpoint))
        checkpoint = torch.load(checkpoint, map_location="cpu")
    sd = checkpoint['state_dict'] if isinstance(checkpoint, dict) and 'state_dict' in checkpoint else checkpoint
    if clean_module is None:
        clean_module = len(sd) > 0 and all(k.startswith('module.') for k in sd)
    if clean_module:
        new = OrderedDict()
        for k, v in sd.items():
            new[k[7:]] = v                      # remove 'module.' of dataparallel
        sd = new
    model.load_state_dict(sd, strict=strict)
    epoch = checkpoint.get('epoch') if isinstance(checkpoint, dict) else None
    best = checkpoint.get('best_pred') if isinstance(checkpoint, dict) else None
    return epoch, best
