"""Confusion-matrix evaluator — B200 drop-in for `utils/metrics.py::Evaluator` (:4-52).

`_generate_matrix` is one atomics-free histogram kernel pass over the int64 gt / pred maps
(16 B/pixel, HBM-bound) that returns the exact int64 [nc, nc] matrix; the 19×19 post-processing
(mIoU etc.) stays in torch on the tiny matrix, like the reference.  The accumulator keeps the
reference's float32 dtype (exact below 2^24 per cell); `confusion_matrix_int64`
additionally keeps the exact integer running sum."""
from __future__ import annotations

import ctypes

import torch

from . import runtime as rt
from ._lib import lib, check


class Evaluator(object):
    def __init__(self, num_class):
        self.num_class = num_class
        self.confusion_matrix = torch.zeros((self.num_class,) * 2)
        self.confusion_matrix_int64 = torch.zeros((self.num_class,) * 2, dtype=torch.int64)

    def Pixel_Accuracy(self):
        return torch.diag(self.confusion_matrix).sum() / self.confusion_matrix.sum()

    def Pixel_Accuracy_Class(self):
        acc = torch.diag(self.confusion_matrix) / self.confusion_matrix.sum(axis=1)
        return self.torch_nanmean(acc)

    def _iou(self):
        cm = self.confusion_matrix
        return torch.diag(cm) / (torch.sum(cm, axis=1) + torch.sum(cm, axis=0) - torch.diag(cm))

    def Mean_Intersection_over_Union(self):
        return self.torch_nanmean(self._iou()).item()

    def Frequency_Weighted_Intersection_over_Union(self):
        freq = torch.sum(self.confusion_matrix, axis=1) / torch.sum(self.confusion_matrix)
        iu = self._iou()
        return (freq[freq > 0] * iu[freq > 0]).sum()

    def _generate_matrix(self, gt_image, pre_image):
        """utils/metrics.py:34-39 — int64 [nc, nc]; rows = gt, cols = pred; gt outside [0, nc) ignored."""
        rt.require_cuda(gt_image, "gt_image")
        rt.require_cuda(pre_image, "pre_image")
        gt = gt_image.to(torch.int64).contiguous()
        pr = pre_image.to(torch.int64).contiguous()
        n_pix = gt.numel()
        out = torch.empty((self.num_class, self.num_class), dtype=torch.int64, device=gt.device)
        nbytes = lib.add_confusion_workspace_bytes(n_pix, self.num_class)
        if nbytes < 0:
            check(int(nbytes), "confusion_workspace_bytes")
        ws = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=gt.device)
        s = ctypes.c_void_p(torch.cuda.current_stream(gt.device).cuda_stream)
        check(lib.add_confusion_matrix(gt.data_ptr(), pr.data_ptr(), n_pix, self.num_class, out.data_ptr(),
                                       ws.data_ptr(), ws.numel(), s), "confusion_matrix")
        return out

    def add_batch(self, gt_image, pre_image):
        assert gt_image.shape == pre_image.shape
        self.add_matrix(self._generate_matrix(gt_image, pre_image))

    def add_matrix(self, cm_int64):
        """Accumulate an int64 confusion matrix produced by the fused head (ADD.evaluate)."""
        if self.confusion_matrix.device != cm_int64.device:
            self.confusion_matrix = self.confusion_matrix.to(cm_int64.device)
            self.confusion_matrix_int64 = self.confusion_matrix_int64.to(cm_int64.device)
        self.confusion_matrix += cm_int64          # fp32 += int64, as in the reference (Q6)
        self.confusion_matrix_int64 += cm_int64

    def all_reduce(self, group=None):
        """Global confusion matrix over the batch shards of all ranks: one all-reduce of
        the exact int64 matrix (2.9 KB); the fp32 accumulator is rebuilt from it.  No-op on one rank."""
        from .parallel import all_reduce_confusion
        all_reduce_confusion(self.confusion_matrix_int64, group)
        self.confusion_matrix = self.confusion_matrix_int64.to(torch.float32)
        return self

    def reset(self):
        self.confusion_matrix = torch.zeros((self.num_class,) * 2).cuda()
        self.confusion_matrix_int64 = torch.zeros((self.num_class,) * 2, dtype=torch.int64).cuda()

    def torch_nanmean(self, x):
        num = torch.where(torch.isnan(x), torch.full_like(x, 0), torch.full_like(x, 1)).sum()
        value = torch.where(torch.isnan(x), torch.full_like(x, 0), x).sum()
        return value / num
