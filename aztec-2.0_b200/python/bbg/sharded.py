"""Multi-GPU plumbing for the MSM: one process per GPU, shard by contiguous point range, combine the partials.

This is the reference's own sharding hook -- Pippenger::pippenger_unsafe(scalars, from, range)
(bb/ecc/curves/bn254/scalar_multiplication/pippenger.cpp:27-31) + g1_sum (c_bind.cpp:40-45) -- spread over
torch.distributed ranks.  The data path has exactly one exchange: an all-gather of one 96-byte Jacobian
partial per rank (NCCL has no elliptic-curve reduction operator), followed by a local g1 sum on every rank.

The compute callables are injected so the host logic can be exercised on CPU (gloo, world_size 2) in tests/ with the
oracle standing in for the device; in the product they are bbg.Pippenger.pippenger_unsafe and bbg.g1_sum.
"""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous [from, from + range) of rank `rank` out of `world`; the first n % world ranks get one extra point."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, base + (1 if rank < extra else 0)


def all_gather_partials(partial, world, device=None):
    """partial: numpy uint64[12] (96 B).  Returns numpy uint64[world, 12] in rank order."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(partial, dtype=np.uint64).view(np.uint8).copy())
    if device is not None:
        t = t.to(device)
    out = torch.empty((world, 96), dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t) if t.is_cuda else dist.all_gather(list(out.unbind(0)), t)
    return out.cpu().numpy().view(np.uint64).reshape(world, 12)


def msm_sharded(partial_fn, sum_fn, scalars, n_total, rank, world, device=None):
    """Every rank passes the scalars of ITS range (`shard_range(n_total, rank, world)`); returns the full MSM
    (identical on every rank).

    partial_fn(scalars_of_range, from, range) -> 96-byte Jacobian partial   (bbg.Pippenger.pippenger_unsafe)
    sum_fn(partials[world, 12])               -> 96-byte Jacobian            (bbg.g1_sum)
    """
    lo, cnt = shard_range(n_total, rank, world)
    scalars = np.asarray(scalars, dtype=np.uint64).reshape(-1, 4)
    if scalars.shape[0] != cnt:
        raise ValueError("rank %d must pass %d scalars, got %d" % (rank, cnt, scalars.shape[0]))
    partial = partial_fn(scalars, lo, cnt)
    if world == 1:
        return np.asarray(partial, dtype=np.uint64).reshape(12)
    return sum_fn(all_gather_partials(partial, world, device))
