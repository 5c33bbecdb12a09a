"""SynchronizedBatchNorm2d, training mode.

CPU: the oracle restatement against fixtures made by the unmodified reference class (tests/golden/syncbn.npz), and
the N>1 host logic over gloo with world_size 2 (pack -> ONE all-reduce -> the reference's formulas == the oracle over
all shards).  GPU: the CUDA kernels (add_bn_stats_fwd / add_bn_finalize / add_bn_apply_fwd behind the drop-in module)
against the same fixtures.  Tolerances: fp32 max-norm relative 1e-5 on outputs and running statistics (sums are
accumulated in a different order); bf16 activations 2^-7."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util
import add_b200
from add_b200 import sync_batchnorm as sbn
from util import orc

GOLD = np.load(util.ROOT / "tests/golden/syncbn.npz")
CASES = list(util.SYNCBN_CASES)


def _params(state):
    return state.get("weight"), state.get("bias"), state["running_mean"], state["running_var"]


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("mode", ["sync", "local"])
def test_oracle_matches_reference_fixture(name, mode):
    spec = util.SYNCBN_CASES[name]
    shards, state = util.make_syncbn_case(name)
    w, b, rm, rv = _params(state)
    outs, nrm, nrv, mean, inv_std = orc.sync_batchnorm_train(shards, w, b, rm, rv, spec["momentum"], spec["eps"], mode == "sync")
    for i, o in enumerate(outs):
        assert util.rel_err(o, torch.from_numpy(GOLD[f"{name}/{mode}/y{i}"])) < 2e-6
    assert util.rel_err(nrm, torch.from_numpy(GOLD[f"{name}/{mode}/running_mean"])) < 2e-6
    assert util.rel_err(nrv, torch.from_numpy(GOLD[f"{name}/{mode}/running_var"])) < 2e-6
    if mode == "sync":
        assert util.rel_err(mean, torch.from_numpy(GOLD[f"{name}/sync/mean"])) < 2e-6
        assert util.rel_err(inv_std, torch.from_numpy(GOLD[f"{name}/sync/inv_std"])) < 2e-6


def test_sync_and_local_formulas_differ_below_eps():
    """clamp(var, eps) vs var + eps: the fixture with per-channel variance below eps separates the two paths."""
    a, b = GOLD["c12_tiny_var/sync/y0"], GOLD["c12_tiny_var/local/y0"]
    assert np.abs(a - b).max() > 1e-2 * np.abs(a).max()


# ---- N > 1 host logic on CPU (gloo, world 2) -------------------------------------------------------
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, name, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shards, state = util.make_syncbn_case(name)
    x = shards[rank]
    C = x.shape[1]
    f = x.reshape(x.shape[0], C, -1)
    local = torch.cat([f.sum(dim=(0, 2)), (f ** 2).sum(dim=(0, 2))])      # what add_bn_stats_fwd produces per rank
    packed = sbn.pack_stats(local, f.shape[0] * f.shape[2])
    sbn.reduce_stats(packed)                                              # the ONE collective of the layer
    if rank == 0:
        torch.save(packed, out_path)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world2_packed_allreduce_gives_the_global_statistics(tmp_path):
    name = "c8_2shards"
    spec = util.SYNCBN_CASES[name]
    out = tmp_path / "packed.pt"
    mp.spawn(_worker, args=(2, _free_port(), name, str(out)), nprocs=2, join=True)
    packed = torch.load(out)
    C = spec["C"]
    n = float(packed[-1])
    assert n == sum(s[0] * s[1] * s[2] for s in spec["shards"])
    mean = packed[:C] / n
    sumvar = packed[C:2 * C] - packed[:C] * mean
    inv_std = (sumvar / n).clamp(spec["eps"]) ** -0.5
    assert util.rel_err(mean, torch.from_numpy(GOLD[f"{name}/sync/mean"])) < 1e-5
    assert util.rel_err(inv_std, torch.from_numpy(GOLD[f"{name}/sync/inv_std"])) < 1e-5


def _worker_bwd(rank, world, port, name, out_path):
    """Backward of one SynchronizedBatchNorm2d layer on rank `rank`: the forward statistics through the packed exchange,
    then the backward's ONE exchange of [sum dy | sum dy*xhat] (csrc/backward.cu: add_bn_bwd_reduce -> exchange_sum ->
    add_bn_bwd_apply; here the per-rank sums are torch, the exchange is the library's gloo path) and
    dx = gamma * inv_std * (dy - sum1/M - xhat * sum2/M)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shards, state = util.make_syncbn_case(name)
    spec = util.SYNCBN_CASES[name]
    x = shards[rank]
    C = x.shape[1]
    g = torch.Generator().manual_seed(900 + rank)
    dy = torch.randn(x.shape, generator=g)
    f = x.reshape(x.shape[0], C, -1)
    packed = sbn.pack_stats(torch.cat([f.sum(dim=(0, 2)), (f ** 2).sum(dim=(0, 2))]), f.shape[0] * f.shape[2])
    sbn.exchange_sum(packed)                                              # forward: [sum | ssum | n]
    M = float(packed[-1])
    mean = packed[:C] / M
    inv_std = ((packed[C:2 * C] - packed[:C] * mean) / M).clamp(spec["eps"]) ** -0.5
    xhat = (f - mean.view(1, C, 1)) * inv_std.view(1, C, 1)
    d = dy.reshape(x.shape[0], C, -1)
    sums = torch.cat([d.sum(dim=(0, 2)), (d * xhat).sum(dim=(0, 2))])
    sbn.exchange_sum(sums)                                                # backward: [sum dy | sum dy*xhat]
    gamma = state["weight"] if state.get("weight") is not None else torch.ones(C)
    dx = (gamma * inv_std).view(1, C, 1) * (d - (sums[:C] / M).view(1, C, 1) - xhat * (sums[C:] / M).view(1, C, 1))
    torch.save(dict(dx=dx.view(x.shape), dgamma=sums[C:].clone(), dbeta=sums[:C].clone()), f"{out_path}.{rank}")
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_world2_backward_exchange_equals_autograd_on_the_whole_batch(tmp_path):
    """The backward direction of the N > 1 path on CPU (gloo, world 2): two ranks, each with its shard and its cotangent,
    exchange ONE [sum dy | sum dy*xhat] vector; every rank's dx, and dgamma / dbeta, equal torch autograd through the
    reference's synchronised formulas (oracle `sync_batchnorm_train`) applied to both shards in one process."""
    name = "c8_2shards"
    out = tmp_path / "bwd"
    mp.spawn(_worker_bwd, args=(2, _free_port(), name, str(out)), nprocs=2, join=True)
    got = [torch.load(f"{out}.{r}") for r in range(2)]
    shards, state = util.make_syncbn_case(name)
    spec = util.SYNCBN_CASES[name]
    xs = [s.clone().requires_grad_(True) for s in shards[:2]]
    w = (state["weight"].clone() if state.get("weight") is not None else torch.ones(spec["C"])).requires_grad_(True)
    b = (state["bias"].clone() if state.get("bias") is not None else torch.zeros(spec["C"])).requires_grad_(True)
    outs, *_ = orc.sync_batchnorm_train(xs, w, b, state["running_mean"], state["running_var"], 0.1, spec["eps"], sync=True)
    dys = [torch.randn(x.shape, generator=torch.Generator().manual_seed(900 + r)) for r, x in enumerate(xs)]
    sum((o * dy).sum() for o, dy in zip(outs, dys)).backward()
    for r in range(2):
        assert util.rel_err(got[r]["dx"], xs[r].grad) < 1e-4, r
        assert util.rel_err(got[r]["dgamma"], w.grad) < 1e-4 and util.rel_err(got[r]["dbeta"], b.grad) < 1e-4


def test_pack_layout():
    p = sbn.pack_stats(torch.arange(6, dtype=torch.float32), 35)
    assert p.tolist() == [0, 1, 2, 3, 4, 5, 35]


def test_state_dict_keys_match_reference_class():
    bn = add_b200.SynchronizedBatchNorm2d(8)
    assert list(bn.state_dict()) == ["weight", "bias", "running_mean", "running_var", "num_batches_tracked"]
    with pytest.raises(RuntimeError):
        bn(torch.zeros(1, 8, 2, 2))          # CPU tensor: no fallback


# ---- CUDA kernels ------------------------------------------------------------------------------------
DEV = "cuda:0"


def _module(name, dev):
    spec = util.SYNCBN_CASES[name]
    _, state = util.make_syncbn_case(name)
    bn = add_b200.SynchronizedBatchNorm2d(spec["C"], eps=spec["eps"], momentum=spec["momentum"], affine=spec["affine"])
    bn.load_state_dict(state, strict=True)
    return bn.to(dev)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_local_batch_statistics(name):
    """One device, no process group: F.batch_norm(training=True) semantics, shard after shard (batchnorm.py:50-53)."""
    shards, _ = util.make_syncbn_case(name)
    bn = _module(name, DEV).train()
    # the kernels use the reference's own sum / square-sum formula (batchnorm.py:116-119) on both paths; ATen's
    # F.batch_norm computes the variance around the mean instead.  With |mean| ~ 1 and std ~ 1e-2 (c12_tiny_var) the
    # fp32 cancellation in ssum - sum * mean shows up at ~3e-5 of the output (eps = 1e-2 dominates the variance there)
    tol = 1e-4 if name == "c12_tiny_var" else 1e-5
    for i, x in enumerate(shards):
        y = bn(x.to(DEV))
        assert util.rel_err(y, torch.from_numpy(GOLD[f"{name}/local/y{i}"])) < tol
    assert util.rel_err(bn.running_mean, torch.from_numpy(GOLD[f"{name}/local/running_mean"])) < 1e-5
    assert util.rel_err(bn.running_var, torch.from_numpy(GOLD[f"{name}/local/running_var"])) < tol
    assert int(bn.num_batches_tracked) == len(shards)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_synchronised_formulas(name):
    """All shards on one GPU as ONE batch with force_sync: sums over the concatenation = sums over the devices, so the
    synchronised path (clamp(var, eps), unbiased running variance over the global count) must reproduce the fixture."""
    shards, _ = util.make_syncbn_case(name)
    if len({tuple(s.shape[2:]) for s in shards}) != 1:
        pytest.skip("shards of different spatial size cannot be concatenated")
    bn = _module(name, DEV).train()
    bn.force_sync = True
    y = bn(torch.cat(shards).to(DEV))
    off = 0
    for i, s in enumerate(shards):
        assert util.rel_err(y[off:off + s.shape[0]], torch.from_numpy(GOLD[f"{name}/sync/y{i}"])) < 1e-5
        off += s.shape[0]
    assert util.rel_err(bn.running_mean, torch.from_numpy(GOLD[f"{name}/sync/running_mean"])) < 1e-5
    assert util.rel_err(bn.running_var, torch.from_numpy(GOLD[f"{name}/sync/running_var"])) < 1e-5


@pytest.mark.gpu
def test_gpu_eval_mode_and_bf16_and_relu():
    name = "c40_3shards"
    shards, state = util.make_syncbn_case(name)
    x = torch.cat(shards)
    bn = _module(name, DEV).eval()
    ref = torch.nn.functional.batch_norm(x, state["running_mean"], state["running_var"], state["weight"], state["bias"],
                                         False, 0.1, 1e-5)
    assert util.rel_err(bn(x.to(DEV)), ref) < 1e-5
    assert util.rel_err(bn(x.to(DEV), relu=True), torch.relu(ref)) < 1e-5
    yb = bn(x.to(DEV).to(torch.bfloat16))
    assert yb.dtype == torch.bfloat16
    refb = torch.nn.functional.batch_norm(x.to(torch.bfloat16).float(), state["running_mean"], state["running_var"],
                                          state["weight"], state["bias"], False, 0.1, 1e-5)
    assert util.rel_err(yb.float(), refb) < 2 ** -7


@pytest.mark.gpu
def test_gpu_statistics_at_training_crop_size():
    """Config-3 shape (769x769 crop, 2 images, 40 channels): per-channel sums against float64 torch sums."""
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(2, 40, 769, 769, generator=g) * 3 + 1).to(DEV).contiguous(memory_format=torch.channels_last)
    bn = add_b200.SynchronizedBatchNorm2d(40).to(DEV).train()
    y = bn(x)
    xd = x.double()
    mean = xd.mean(dim=(0, 2, 3))
    var = xd.var(dim=(0, 2, 3), unbiased=False)
    ref = (xd - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + 1e-5)
    assert util.rel_err(y, ref.float()) < 1e-4
    n = x.numel() / 40
    assert util.rel_err(bn.running_var, (0.9 + 0.1 * var * n / (n - 1)).float()) < 1e-4
