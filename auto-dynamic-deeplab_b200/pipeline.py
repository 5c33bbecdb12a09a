"""Host-fed evaluation loop — what `eval.py::Evaluation.validation` / `.dynamic_inference`
(eval.py:165-230) do around the model: for every batch from the loader copy image + label to the
device, run the network, argmax, add to the confusion matrix.  Here the loop is double-buffered:
the H2D copy of batch i+1 runs on a side stream while batch i computes, and only the per-image
confusion matrices (N x 19 x 19 int64) come back over PCIe.  PyTorch supplies pinned memory,
streams and events; all compute is libadd_b200."""
from __future__ import annotations

from typing import Iterable, Iterator, List, Optional, Tuple

import torch

import ctypes

from . import runtime as rt
from ._lib import lib, check


class HostPipeline:
    """evaluate(batches) yields, per batch of HOST tensors (x fp32 [N,3,H,W], gt int64 or uint8 [N,H,W], ideally
    pinned), `(cm int64 [n_exits or 1, N, nc, nc] pinned host tensor, exit flags or None)`.

    edm=None  → multi-exit `ADD.evaluate` (eval.py:165-193, every exit scored);
    edm given → EDM-gated early exit per image, `ADD.dynamic_evaluate` (eval.py:195-221)."""

    def __init__(self, net, edm=None, threshold: float = 1.0, exit_mode: str = "reference", depth: int = 2):
        self.net, self.edm, self.threshold, self.exit_mode = net, edm, float(threshold), exit_mode
        self.depth = max(2, int(depth))
        self.device = next(net.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("HostPipeline needs the model on a CUDA device (add_b200 has no CPU fallback)")
        self.copy_stream = torch.cuda.Stream(self.device)
        self._slots: Optional[List[dict]] = None
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def _ensure_slots(self, x: torch.Tensor, gt: torch.Tensor) -> None:
        if self._slots is not None and self._slots[0]["x"].shape == x.shape:
            return
        self._slots = []
        for _ in range(self.depth):
            self._slots.append(dict(x=torch.empty(x.shape, dtype=torch.float32, device=self.device),
                                    gt=torch.empty(gt.shape, dtype=torch.int64, device=self.device),
                                    gt_u8=(torch.empty(gt.shape, dtype=torch.uint8, device=self.device)
                                           if gt.dtype == torch.uint8 else None),
                                    ready=torch.cuda.Event(), free=torch.cuda.Event(), out=None))

    def _prefetch(self, slot: dict, x: torch.Tensor, gt: torch.Tensor) -> None:
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(slot["free"])        # the compute that last read this slot is done
            slot["x"].copy_(x, non_blocking=True)
            if gt.dtype == torch.uint8:
                # labels travel at their native 1 byte per pixel and are widened on the device (add_widen_labels_u8)
                if slot["gt_u8"] is None:
                    slot["gt_u8"] = torch.empty(gt.shape, dtype=torch.uint8, device=self.device)
                slot["gt_u8"].copy_(gt, non_blocking=True)
                check(lib.add_widen_labels_u8(slot["gt_u8"].data_ptr(), slot["gt"].data_ptr(), gt.numel(),
                                              ctypes.c_void_p(self.copy_stream.cuda_stream)), "widen_labels_u8")
            else:
                slot["gt"].copy_(gt, non_blocking=True)
            slot["ready"].record(self.copy_stream)
        self.h2d_bytes += x.numel() * x.element_size() + gt.numel() * gt.element_size()

    def evaluate(self, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]]) -> Iterator[Tuple[torch.Tensor, Optional[list]]]:
        main = torch.cuda.current_stream(self.device)
        it = iter(batches)
        nxt = next(it, None)
        if nxt is None:
            return
        self._ensure_slots(*nxt)
        for s in self._slots:
            s["free"].record(main)
        i = 0
        self._prefetch(self._slots[0], *nxt)
        while nxt is not None:
            slot = self._slots[i % self.depth]
            nxt = next(it, None)
            if nxt is not None:                              # enqueue the next copy BEFORE this batch's compute
                self._prefetch(self._slots[(i + 1) % self.depth], *nxt)
            main.wait_event(slot["ready"])
            if self.edm is None:
                cm, flags = self.net.evaluate(slot["x"], slot["gt"]), None
            else:
                # the slot buffers are stable: the plans are recorded directly on them (no device-to-device input copy)
                cm, flags, _ = self.net.dynamic_evaluate(slot["x"], slot["gt"], self.threshold, self.edm, self.exit_mode,
                                                         bind_inputs=True)
                cm = cm.unsqueeze(0)
            slot["free"].record(main)
            if slot["out"] is None or slot["out"].shape != cm.shape:
                slot["out"] = torch.empty(cm.shape, dtype=torch.int64).pin_memory()
            slot["out"].copy_(cm, non_blocking=True)
            main.synchronize()                               # the step's result is on the host
            self.d2h_bytes += cm.numel() * 8
            yield slot["out"], flags
            i += 1
