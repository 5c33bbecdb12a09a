"""SynchronizedBatchNorm2d — drop-in for modeling/sync_batchnorm/batchnorm.py:180 with a B200 forward.

Eval mode inside the ADD plans never reaches this module's `forward`: BN is folded into the neighbouring conv.
Called on its own (and in training mode) it runs three libadd_b200 kernels:

    add_bn_stats_fwd   per-channel [sum, square-sum] of this rank's shard        (batchnorm.py:59-61)
    all_reduce         ONE collective of the packed [sum | ssum | count] vector  (batchnorm.py:90-111: the reference
                       does ReduceAddCoalesced + Broadcast between DataParallel threads; with one process per GPU it
                       is a single NCCL all-reduce over NVLink per BN layer — gloo in the CPU tests)
    add_bn_finalize    mean, inv_std, running-statistics update                  (batchnorm.py:113-125)
    add_bn_apply_fwd   y = (x - mean) * (inv_std * weight) + bias                (batchnorm.py:68-75)

Semantics follow the reference exactly: with a process group of more than one rank (or `force_sync=True`) the
statistics are synchronised and inv_std = clamp(var, eps)^-1/2; otherwise it is `F.batch_norm` with local batch
statistics (inv_std = (var + eps)^-1/2) — what the reference does on one device and under DDP.
No autograd: backward is the remaining part of the training row."""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from ._lib import lib, check, AddTensor, ADD_F32, ADD_BF16, RELU_OUT


def pack_stats(sums: torch.Tensor, count: int) -> torch.Tensor:
    """[sum(C) | ssum(C)] + element count -> the fp32 vector that is all-reduced (2C+1 floats)."""
    out = torch.empty(sums.numel() + 1, dtype=torch.float32, device=sums.device)
    out[:-1] = sums
    out[-1] = float(count)
    return out


def reduce_stats(packed: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Sum the packed statistics over all ranks, in place (no-op without an initialised group of > 1 ranks)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
    return packed


class PeerExchange:
    """The one-kernel all-reduce of csrc/peer.cu for one process group: a symmetric buffer (two slots) + signal pad from
    torch.distributed._symmetric_memory, mapped into every peer; `all_reduce_(vec)` launches ONE kernel on the current
    stream.  NVLink-connected GPUs, one rank per GPU."""
    CAP = 64 * 1024                     # bytes per slot: vectors are 2C+1 floats / 2C doubles (C <= 1280 on this network)
    _instances: dict = {}

    def __init__(self, group=None):
        import numpy as np
        import torch.distributed._symmetric_memory as symm_mem
        group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.buf = symm_mem.empty(2 * self.CAP, dtype=torch.uint8, device=dev)
        self.hdl = symm_mem.rendezvous(self.buf, group)
        self.buf_ptrs = np.array([int(p) for p in self.hdl.buffer_ptrs], dtype=np.uint64)
        self.sig_ptrs = np.array([int(p) for p in self.hdl.signal_pad_ptrs], dtype=np.uint64)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        self.epoch_dev = torch.zeros(1, dtype=torch.int32, device=dev)     # call counter, incremented by the kernel (graph-replayable)
        self.calls = 0

    @classmethod
    def get(cls, group=None) -> "PeerExchange":
        key = id(group) if group is not None else 0
        if key not in cls._instances:
            cls._instances[key] = cls(group)
        return cls._instances[key]

    def all_reduce_(self, vec: torch.Tensor) -> torch.Tensor:
        assert vec.is_cuda and vec.is_contiguous() and vec.dtype in (torch.float32, torch.float64)
        self.calls += 1
        s = ctypes.c_void_p(torch.cuda.current_stream(vec.device).cuda_stream)
        check(lib.add_peer_allreduce(vec.data_ptr(), vec.numel(), vec.element_size(), self.buf_ptrs.ctypes.data, self.sig_ptrs.ctypes.data,
                                     self.rank, self.world, 0, self.epoch_dev.data_ptr(), self.CAP, self.status.data_ptr(), s),
              "peer_allreduce")
        return vec

    def check_status(self) -> None:
        """Host-side check (a device sync): raises if any exchange timed out waiting for a peer."""
        if int(self.status.item()) != 0:
            raise RuntimeError("add_b200 peer exchange: a peer rank did not arrive (timed out)")


_EXCHANGE = {"mode": "auto"}     # 'auto': peer-memory kernel on NCCL groups of CUDA ranks, dist.all_reduce otherwise; 'nccl'; 'peer'


def set_exchange_mode(mode: str) -> None:
    assert mode in ("auto", "nccl", "peer")
    _EXCHANGE["mode"] = mode


def exchange_sum(vec: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """vec <- sum over the ranks, in place: the per-layer SynchronizedBatchNorm2d exchange (forward [sum | ssum | n],
    backward [sum dy | sum dy*xhat]).  On GPUs it is ONE peer-memory kernel (PeerExchange); on the CPU / gloo tests and as
    a fallback when symmetric memory cannot be set up it is `dist.all_reduce`."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) <= 1:
        return vec
    mode = _EXCHANGE["mode"]
    if vec.is_cuda and mode in ("auto", "peer") and dist.get_backend(group) == "nccl":
        try:
            return PeerExchange.get(group).all_reduce_(vec)
        except Exception:
            if mode == "peer":
                raise
            _EXCHANGE["mode"] = "nccl"          # symmetric memory unavailable on this system: fall back for good
    dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return vec


def _desc(t: torch.Tensor) -> AddTensor:
    """NHWC view descriptor of a logical-NCHW channels_last tensor."""
    n, c, h, w = t.shape
    return AddTensor(t.data_ptr(), n, h, w, c, c, ADD_BF16 if t.dtype == torch.bfloat16 else ADD_F32)


def batch_norm_forward(bn: nn.BatchNorm2d, x: torch.Tensor, relu: bool = False, sync: Optional[bool] = None,
                       group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """BatchNorm2d forward of any `_BatchNorm` parameter holder through libadd_b200 (no ATen batch_norm): training mode
    = batch statistics (+ running-statistics update), eval mode = running statistics; optional fused ReLU.
    sync: True = statistics all-reduced over `group`, inv_std = clamp(var, eps)^-1/2 (batchnorm.py:113-125);
    False = this rank's statistics, inv_std = (var + eps)^-1/2 (F.batch_norm, batchnorm.py:50-53);
    None = synchronised iff a process group with more than one rank is initialised."""
    if not x.is_cuda:
        raise RuntimeError("add_b200 runs on CUDA tensors only (no CPU fallback)")
    if x.dim() != 4 or x.shape[1] != bn.num_features:
        raise ValueError(f"expected [N, {bn.num_features}, H, W], got {tuple(x.shape)}")
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    x = x.contiguous(memory_format=torch.channels_last)          # zero copy when already NHWC
    n, c, h, w = x.shape
    stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    xd = _desc(x)
    w_ptr = bn.weight.data_ptr() if bn.affine else None
    b_ptr = bn.bias.data_ptr() if bn.affine else None
    stats = torch.empty(2 * c, dtype=torch.float32, device=x.device)     # [mean | inv_std]
    mean, inv_std = stats[:c], stats[c:]
    if bn.training or not bn.track_running_stats:
        ws_bytes = lib.add_bn_stats_workspace_bytes(n, h, w, c)
        check(ws_bytes if ws_bytes < 0 else 0, "bn_stats_workspace_bytes")
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
        packed = torch.empty(2 * c + 1, dtype=torch.float32, device=x.device)
        check(lib.add_bn_stats_fwd(ctypes.byref(xd), packed.data_ptr(), ws.data_ptr(), ws_bytes, stream), "bn_stats")
        packed[-1:].fill_(float(n * h * w))      # (a Python-scalar setitem is a host-to-device copy: not graph-capturable)
        if sync is None:
            sync = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
        if sync:
            exchange_sum(packed, group)
        momentum = 0.0
        rm = rv = None
        if bn.training and bn.track_running_stats:
            if bn.num_batches_tracked is not None:
                bn.num_batches_tracked += 1
            momentum = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
            rm, rv = bn.running_mean.data_ptr(), bn.running_var.data_ptr()
        check(lib.add_bn_finalize(packed.data_ptr(), packed.data_ptr() + 8 * c, 0.0, c, float(bn.eps), float(momentum),
                                  1 if sync else 0, rm, rv, mean.data_ptr(), inv_std.data_ptr(), stream), "bn_finalize")
    else:
        # eval: F.batch_norm with the running statistics (batchnorm.py:50-53); C-length vectors, plumbing
        mean.copy_(bn.running_mean)
        torch.rsqrt(bn.running_var + bn.eps, out=inv_std)
    y = torch.empty_like(x)                                      # keeps channels_last
    yd = _desc(y)
    check(lib.add_bn_apply_fwd(ctypes.byref(xd), ctypes.byref(yd), mean.data_ptr(), inv_std.data_ptr(), w_ptr, b_ptr,
                               RELU_OUT if relu else 0, stream), "bn_apply")
    return y


class SynchronizedBatchNorm2d(nn.BatchNorm2d):
    """Same constructor, parameters and state_dict keys as the reference class (a `_BatchNorm` subclass)."""

    def __init__(self, num_features, eps=1e-5, momentum=0.1, affine=True, process_group=None, force_sync=False):
        super().__init__(num_features, eps=eps, momentum=momentum, affine=affine)
        self.process_group = process_group
        self.force_sync = force_sync        # take the synchronised formulas even with one rank (tests)

    def forward(self, x: torch.Tensor, relu: bool = False) -> torch.Tensor:
        return batch_norm_forward(self, x, relu, True if self.force_sync else None, self.process_group)
