"""Host-fed evaluation loop — what `eval.py::Evaluation.validation` / `.dynamic_inference`
(eval.py:165-230) do around the model: for every batch from the loader copy image + label to the
device, run the network, argmax, add to the confusion matrix.  Here the loop runs over three slots:
the H2D copy of batch i+2 (copy stream) and the trunk of batch i+1 are in flight while the host
decides batch i's exits, and only the per-image confusion matrices (N x 19 x 19 int64) come back
over PCIe.  Images may arrive as the loader's fp32 NCHW tensors or as uint8 HWC (normalised on the
device), labels as int64 or uint8.  PyTorch supplies pinned memory, streams and events; all
compute is libadd_b200."""
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

    CITYSCAPES_NORM = ((0.29866842, 0.30135223, 0.30561872), (0.23925215, 0.23859318, 0.2385942))   # cityscapes.py:53-54

    def __init__(self, net, edm=None, threshold: float = 1.0, exit_mode: str = "reference", depth: int = 3,
                 image_norm=CITYSCAPES_NORM):
        """image_norm = (mean, std): used when a batch's images arrive as uint8 [N,H,W,3] (the PNG bytes): they cross
        PCIe at 3 bytes per pixel and are normalised on the device exactly as the reference's host transforms do."""
        self.net, self.edm, self.threshold, self.exit_mode = net, edm, float(threshold), exit_mode
        self.image_norm = image_norm
        # 3 slots: batch i+2 is being copied in while the trunk of batch i+1 and the exit heads of batch i compute
        self.depth = max(2 if edm is None else 3, int(depth))
        self.device = next(net.parameters()).device
        rt.require_cuda_device(self.device, "HostPipeline")
        self.copy_stream = torch.cuda.Stream(self.device)
        self._slots: Optional[List[dict]] = None
        self._slot_sets: dict = {}
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    @staticmethod
    def _batch_key(batch) -> tuple:
        x, gt = batch
        return (tuple(x.shape), x.dtype, tuple(gt.shape), gt.dtype)

    def _ensure_slots(self, x: torch.Tensor, gt: torch.Tensor) -> None:
        """Slots (device buffers + the plans recorded on them) are per batch SHAPE: a loader's last, smaller batch
        (500 Cityscapes val images in batches of 8 leave 4) gets its own slot set instead of being broadcast into —
        and counted as — a full one."""
        shape = (x.shape[0], 3, x.shape[1], x.shape[2]) if x.dtype == torch.uint8 else tuple(x.shape)
        if gt.shape[0] != shape[0] or tuple(gt.shape[1:]) != tuple(shape[2:]):
            raise ValueError(f"labels {tuple(gt.shape)} do not match images {shape}")
        key = (shape, tuple(gt.shape), gt.dtype)
        if key in self._slot_sets:
            self._slots = self._slot_sets[key]
            return
        self._slots = self._slot_sets[key] = []
        for _ in range(self.depth):
            self._slots.append(dict(x=torch.empty(shape, dtype=torch.float32, device=self.device), x_u8=None,
                                    gt=torch.empty(gt.shape, dtype=torch.int64, device=self.device),
                                    gt_u8=(torch.empty(gt.shape, dtype=torch.uint8, device=self.device)
                                           if gt.dtype == torch.uint8 else None),
                                    ready=torch.cuda.Event(), free=torch.cuda.Event(), out=None))

    def _prefetch(self, slot: dict, x: torch.Tensor, gt: torch.Tensor) -> None:
        want = (x.shape[0], 3, x.shape[1], x.shape[2]) if x.dtype == torch.uint8 else tuple(x.shape)
        if tuple(slot["x"].shape) != want or tuple(slot["gt"].shape) != tuple(gt.shape):
            raise RuntimeError(f"HostPipeline slot {tuple(slot['x'].shape)} fed a batch of shape {want}")   # never broadcast
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(slot["free"])        # the compute that last read this slot is done
            if x.dtype == torch.uint8:
                # uint8 HWC images: 3 bytes per pixel over PCIe, normalised to fp32 NCHW on the device
                if x.dim() != 4 or x.shape[3] != 3:
                    raise ValueError(f"uint8 images must be [N,H,W,3], got {tuple(x.shape)}")
                if slot["x_u8"] is None or slot["x_u8"].shape != x.shape:
                    slot["x_u8"] = torch.empty(x.shape, dtype=torch.uint8, device=self.device)
                slot["x_u8"].copy_(x, non_blocking=True)
                (m0, m1, m2), (s0, s1, s2) = self.image_norm
                check(lib.add_normalize_u8_hwc_to_nchw(slot["x_u8"].data_ptr(), slot["x"].data_ptr(), x.shape[0], x.shape[1],
                                                       x.shape[2], m0, m1, m2, s0, s1, s2,
                                                       ctypes.c_void_p(self.copy_stream.cuda_stream)), "normalize_u8")
            else:
                slot["x"].copy_(x, non_blocking=True)
            if gt.dtype == torch.uint8:
                # labels travel at their native 1 byte per pixel; the gated path reads them as they are
                # (add_upsample_argmax_u8_fwd), the multi-exit path widens them on the device (add_widen_labels_u8)
                if slot["gt_u8"] is None:
                    slot["gt_u8"] = torch.empty(gt.shape, dtype=torch.uint8, device=self.device)
                slot["gt_u8"].copy_(gt, non_blocking=True)
                if self.edm is None:
                    check(lib.add_widen_labels_u8(slot["gt_u8"].data_ptr(), slot["gt"].data_ptr(), gt.numel(),
                                                  ctypes.c_void_p(self.copy_stream.cuda_stream)), "widen_labels_u8")
            else:
                slot["gt"].copy_(gt, non_blocking=True)
            slot["labels_u8"] = gt.dtype == torch.uint8
            slot["ready"].record(self.copy_stream)
        self.h2d_bytes += x.numel() * x.element_size() + gt.numel() * gt.element_size()

    def evaluate(self, batches: Iterable[Tuple[torch.Tensor, torch.Tensor]]) -> Iterator[Tuple[torch.Tensor, Optional[list]]]:
        """Runs of equal-shaped batches are pipelined; a shape change (the loader's final partial batch) drains the
        pipeline and continues on that shape's own slots."""
        import itertools
        for _, run in itertools.groupby(batches, key=self._batch_key):
            if self.edm is not None:
                yield from self._evaluate_gated(run)
            else:
                yield from self._evaluate_all_exits(run)

    def _evaluate_all_exits(self, batches):
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
            cm, flags = self.net.evaluate(slot["x"], slot["gt"]), None
            slot["free"].record(main)
            if slot["out"] is None or slot["out"].shape != cm.shape:
                slot["out"] = torch.empty(cm.shape, dtype=torch.int64).pin_memory()
            slot["out"].copy_(cm, non_blocking=True)
            main.synchronize()                               # the step's result is on the host
            self.d2h_bytes += cm.numel() * 8
            yield slot["out"], flags
            i += 1

    def _evaluate_gated(self, batches):
        """EDM-gated early exit, software-pipelined over the slots so that neither host round trip of a step leaves the
        GPU idle: in iteration i the trunk of batch i+1 is enqueued (`dynamic_evaluate_begin`) BEFORE the host waits
        for batch i's gate values and launches its exit heads / remaining trunk (`dynamic_evaluate_finish`), and the
        confusion matrices of batch i are awaited one iteration later.  Stream order:
            trunk(i+1) | heads(i), rest(i), D2H cm(i) | trunk(i+2) | heads(i+1) ...   with H2D(i+2) on the copy stream.
        The slot buffers are stable, so the plans are recorded directly on them (no device-to-device input copy)."""
        main = torch.cuda.current_stream(self.device)
        it = iter(batches)
        first = next(it, None)
        if first is None:
            return
        self._ensure_slots(*first)
        for s in self._slots:
            s["free"].record(main)
            s.setdefault("done", torch.cuda.Event())
        D = self.depth

        def begin(slot):
            main.wait_event(slot["ready"])
            gt = slot["gt_u8"] if slot.get("labels_u8") else slot["gt"]
            return self.net.dynamic_evaluate_begin(slot["x"], gt, self.threshold, self.edm, self.exit_mode,
                                                   bind_inputs=True)

        queued = [first]                                    # host batches whose H2D has been enqueued, not yet begun
        self._prefetch(self._slots[0], *first)
        nb = next(it, None)
        if nb is not None:
            self._prefetch(self._slots[1 % D], *nb)
            queued.append(nb)
        handle = begin(self._slots[0])
        queued.pop(0)
        i = 0
        prev = None                                         # (slot, flags) of the batch whose result is still in flight
        while handle is not None:
            slot = self._slots[i % D]
            nb = next(it, None)
            if nb is not None:                              # batch i+2 -> the slot batch i-1 has left
                self._prefetch(self._slots[(i + 2) % D], *nb)
                queued.append(nb)
            nxt_handle = None
            if queued:                                      # trunk of batch i+1 goes in front of this batch's decision
                nxt_handle = begin(self._slots[(i + 1) % D])
                queued.pop(0)
            cm, flags, _ = self.net.dynamic_evaluate_finish(handle)
            cm = cm.unsqueeze(0)
            slot["free"].record(main)
            if slot["out"] is None or slot["out"].shape != cm.shape:
                slot["out"] = torch.empty(cm.shape, dtype=torch.int64).pin_memory()
            slot["out"].copy_(cm, non_blocking=True)
            slot["done"].record(main)
            self.d2h_bytes += cm.numel() * 8
            if prev is not None:
                prev[0]["done"].synchronize()               # the previous step's result is on the host
                yield prev[0]["out"], prev[1]
            prev = (slot, flags)
            handle = nxt_handle
            i += 1
        if prev is not None:
            prev[0]["done"].synchronize()
            yield prev[0]["out"], prev[1]


class ResidentPipeline:
    """The same software pipeline for batches that already live in HBM: `evaluate(batches)` takes DEVICE tensors
    (x fp32 [N,3,H,W], gt int64 [N,H,W]) that cycle over at least `depth` distinct, stable buffers (the plans are
    recorded on them), and yields `(cm int64 [N,nc,nc] device tensor, exit flags)` one batch behind the enqueue front.
    The yielded cm aliases plan buffers of that slot; it is overwritten `depth` batches later."""

    def __init__(self, net, edm, threshold: float = 1.0, exit_mode: str = "reference"):
        self.net, self.edm, self.threshold, self.exit_mode = net, edm, float(threshold), exit_mode

    def evaluate(self, batches):
        it = iter(batches)
        cur = next(it, None)
        if cur is None:
            return
        handle = self.net.dynamic_evaluate_begin(cur[0], cur[1], self.threshold, self.edm, self.exit_mode, bind_inputs=True)
        while handle is not None:
            nxt = next(it, None)
            nxt_handle = None
            if nxt is not None:
                nxt_handle = self.net.dynamic_evaluate_begin(nxt[0], nxt[1], self.threshold, self.edm, self.exit_mode,
                                                             bind_inputs=True)
            cm, flags, _ = self.net.dynamic_evaluate_finish(handle)
            yield cm, flags
            handle = nxt_handle
