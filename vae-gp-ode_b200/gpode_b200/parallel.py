"""Data-parallel plumbing for the rollout: one process per GPU, trajectories sharded, parameters
replicated, ONE all-reduce of the flat gradient buffer per iteration.

The reference has no distributed code; the path shards because every (trajectory, MC-sample) state
evolves independently given the per-sample parameter set (experiments/model/core/odegpvae.py:41-43).
Every rank must draw the SAME function samples (seed the host RNG helpers identically) so that nu,
omega etc. are replicas; only z0 differs.  Batch-mean losses (experiments/model/create_model.py:53,58)
must be scaled by local_N / global_N before the sum all-reduce; the replicated KL term is counted once.
"""
import torch
import torch.distributed as dist


def shard_bounds(n_items, rank, world):
    """Contiguous block [lo, hi) of rank `rank`; the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_trajectories(z0, rank=None, world=None):
    """Slice the trajectory axis (dim -2) of z0 (N,D) or (L,N,D) for this rank."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    if z0.shape[-2] < world:
        # an empty shard would make this rank raise inside the rollout (N < 1 is a shape error) while the others block in the
        # all-reduce: fail on EVERY rank, before any collective
        raise ValueError("cannot shard %d trajectories over %d ranks: every rank needs at least one" % (z0.shape[-2], world))
    lo, hi = shard_bounds(z0.shape[-2], rank, world)
    return z0[..., lo:hi, :]


def allreduce_gradients(tensors, group=None):
    """Sum-all-reduce a list of gradient tensors as ONE flat buffer (one latency-bound collective over
    NVLink instead of one per tensor) and scatter the result back in place."""
    tensors = [t for t in tensors if t is not None]
    if not tensors or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return tensors
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
    return tensors


def scale_local_mean(value, n_local, n_global):
    """Turn a local batch mean into this rank's share of the global batch mean."""
    return value * (float(n_local) / float(n_global))
