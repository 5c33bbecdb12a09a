"""CPU, world_size 2 over gloo: the N>1 path's host logic — contiguous batch shards, per-rank
confusion matrices (oracle as the stand-in for the per-GPU kernel), one int64 all-reduce — equals
the single-process result bit for bit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util
import add_b200
from util import orc


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, gt, pred, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = add_b200.shard_range(gt.shape[0], world, rank)
    ev = add_b200.Evaluator(19)
    for i in range(lo, hi):      # this rank's images only
        ev.add_matrix(torch.from_numpy(orc.generate_matrix(gt[i:i + 1].numpy(), pred[i:i + 1].numpy())))
    ev.all_reduce()
    if rank == 0:
        torch.save((ev.confusion_matrix_int64, ev.confusion_matrix, ev.Mean_Intersection_over_Union()), out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_is_a_partition():
    for n in (0, 1, 5, 8, 16, 17):
        for world in (1, 2, 3, 4, 8):
            spans = [add_b200.shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        add_b200.shard_range(4, 2, 2)


@pytest.mark.timeout(120)
def test_world2_confusion_allreduce_matches_single_process(tmp_path):
    g = torch.Generator().manual_seed(77)
    gt = torch.randint(0, 19, (5, 24, 31), generator=g)
    gt[torch.rand(5, 24, 31, generator=g) < 0.1] = 255
    pred = torch.randint(0, 19, (5, 24, 31), generator=g)
    out = tmp_path / "r0.pt"
    mp.spawn(_worker, args=(2, _free_port(), gt, pred, str(out)), nprocs=2, join=True)
    cm64, cm32, miou = torch.load(out)
    want = orc.generate_matrix(gt.numpy(), pred.numpy())
    assert np.array_equal(cm64.numpy(), want)
    assert np.array_equal(cm32.numpy(), want.astype(np.float32))
    ev = add_b200.Evaluator(19)
    ev.add_matrix(torch.from_numpy(want))
    assert miou == ev.Mean_Intersection_over_Union()
