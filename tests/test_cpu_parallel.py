"""CPU, gloo, world_size 2: host-side logic of the data-parallel path (sharding + flat gradient all-reduce)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gpode_b200 import parallel as PL


def test_shard_bounds_cover_everything():
    for n in (1, 2, 7, 25, 256, 65536):
        for world in (1, 2, 3, 8):
            spans = [PL.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        z0 = torch.arange(25 * 6, dtype=torch.float32).view(25, 6)
        mine = PL.shard_trajectories(z0)
        # stand-in for the kernel-level gradients of this rank's shard: linear in the shard, so the sum is known
        g = [mine.sum(0), mine.pow(2).sum().reshape(1), torch.full((3, 2), float(rank + 1))]
        PL.allreduce_gradients(g)
        loss = PL.scale_local_mean(mine.mean(), mine.shape[0], z0.shape[0])
        t = torch.tensor([loss])
        dist.all_reduce(t)
        out[rank] = (mine.shape[0], g[0].clone(), g[1].clone(), g[2].clone(), t.item())
    finally:
        dist.destroy_process_group()


def test_gloo_world2_allreduce_and_loss_normalisation():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    z0 = torch.arange(25 * 6, dtype=torch.float32).view(25, 6)
    assert out[0][0] + out[1][0] == 25 and out[0][0] == 13
    for r in range(world):
        assert torch.allclose(out[r][1], z0.sum(0))
        assert torch.allclose(out[r][2], z0.pow(2).sum().reshape(1))
        assert torch.allclose(out[r][3], torch.full((3, 2), 3.0))
        assert abs(out[r][4] - z0.mean().item()) < 1e-4


def test_fewer_trajectories_than_ranks_fails_on_every_rank():
    """an empty shard would raise inside the rollout on some ranks while the others block in the all-reduce: refuse up front, everywhere"""
    import pytest
    z0 = torch.zeros(3, 6)
    for rank in range(4):
        with pytest.raises(ValueError):
            PL.shard_trajectories(z0, rank=rank, world=4)
    assert PL.shard_trajectories(z0, rank=2, world=3).shape == (1, 6)
