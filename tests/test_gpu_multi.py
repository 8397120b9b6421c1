"""GPU, 2 ranks (NCCL): trajectory-sharded rollout + ONE all-reduce of the kernel-level gradients equals the single-GPU gradients.
Skipped on a box with fewer than two GPUs (the host-side logic is covered by tests/test_cpu_parallel.py with gloo)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _problem(device, D=16, M=64, S=48, N=4099, T=4, L=2, seed=0):
    rs = np.random.RandomState(seed)
    f32 = lambda a: torch.tensor(a, dtype=torch.float32, device=device)
    return dict(Z=f32(rs.normal(size=(M, D))), ell=f32(1.5 + rs.uniform(size=(D, D))), var=f32(0.5 + rs.uniform(size=D)),
                nu=f32(0.3 * rs.normal(size=(L, D, M, 1))), eps=f32(rs.normal(size=(L, D, S, D))), phase=f32(rs.uniform(size=(L, 1, S, D)) * 2 * np.pi),
                w=f32(rs.normal(size=(L, S, D))), z0=f32(rs.normal(size=(N, D))), G=f32(rs.normal(size=(L, N, T, D))),
                ts=0.1 * torch.arange(T, dtype=torch.float32, device=device))


def _grads(p, lo, hi):
    import gpode_b200 as gp
    leaves = {k: p[k].clone().requires_grad_(True) for k in ("Z", "nu", "ell", "var")}
    z0 = p["z0"][lo:hi].clone().requires_grad_(True)
    traj = gp.gp_rollout(z0, p["ts"], leaves["Z"], leaves["nu"], p["eps"], p["phase"], p["w"], leaves["ell"], leaves["var"], "rbf_dimwise", 1, "rk4")
    (traj * p["G"][:, lo:hi]).sum().backward()
    return z0.grad, [leaves[k].grad for k in ("Z", "nu", "ell", "var")]


def _worker(rank, world, port, out):
    for q in (ROOT, os.path.join(ROOT, "vae-gp-ode_b200")):
        if q not in sys.path:
            sys.path.insert(0, q)
    import torch.distributed as dist
    from gpode_b200 import parallel as PL
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        p = _problem("cuda:%d" % rank)                      # the same seeded problem on every rank (replicated parameters / samples)
        lo, hi = PL.shard_bounds(p["z0"].shape[0], rank, world)
        dz, pg = _grads(p, lo, hi)
        PL.allreduce_gradients(pg)                          # one flat NCCL all-reduce
        out[rank] = (lo, hi, dz.cpu(), [g.cpu() for g in pg])
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_allreduced_gradients_equal_single_gpu():
    import torch.multiprocessing as mp
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    p = _problem("cuda:0")
    dz_full, pg_full = _grads(p, 0, p["z0"].shape[0])
    rel = lambda a, b: ((a.double() - b.double()).norm() / b.double().norm()).item()
    dz_sharded = torch.cat([out[r][2] for r in range(world)])
    assert out[0][0] == 0 and out[0][1] == out[1][0] and out[1][1] == p["z0"].shape[0]
    assert rel(dz_sharded, dz_full.cpu()) < 1e-6            # per-trajectory quantity: identical kernels on each shard
    for r in range(world):
        for name, a, b in zip(("dZ", "dnu", "dell", "dvar"), out[r][3], pg_full):
            e = rel(a, b.cpu())
            print("rank %d sharded + all-reduced %s vs single GPU: %.2e" % (r, name, e))
            assert e < 1e-5, (name, e)
