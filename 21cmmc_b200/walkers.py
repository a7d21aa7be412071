"""Walker batching and sharding: the N-GPU side of the likelihood (SURVEY.md sections 8e, 8f-1).

Walkers are independent (the reference maps them over a process pool, src/py21cmmc/mcmc.py:580-583,
ensemble.py:204-208).  Here walker ``w`` belongs to rank ``w % world``; every rank evaluates its own
walkers with one batched device call and the only exchange is an all-gather of the per-walker fp64
log-likelihoods (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_indices(nwalkers: int, rank: int, world: int):
    """Walker ids owned by ``rank`` (round robin)."""
    return list(range(rank, nwalkers, world))


def gather_walker_values(local, nwalkers: int, rank: int, world: int, device=None, group=None):
    """All-gather per-walker scalars and return them in walker order on every rank.

    ``local``: 1-D float64 array/tensor with the values of ``shard_indices(nwalkers, rank, world)``.
    """
    import torch
    import torch.distributed as dist

    per = -(-nwalkers // world)
    t = torch.as_tensor(local, dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    buf = torch.full((per,), float("nan"), dtype=torch.float64, device=t.device)
    buf[: t.numel()] = t
    if world == 1:
        return buf[:nwalkers].cpu().numpy()
    out = torch.empty((world * per,), dtype=torch.float64, device=t.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    out = out.view(world, per).cpu().numpy()
    full = np.empty(nwalkers)
    for r in range(world):
        idx = shard_indices(nwalkers, r, world)
        full[idx] = out[r, : len(idx)]
    return full


class BatchedWalkerPool:
    """``pool``-compatible object: ``map(fn, plist)`` evaluates all walkers with batched device calls.

    Pass it where the reference passes its ``ProcessPoolExecutor`` (``mcmc.py:580-592``); ``fn`` is
    ignored in favour of ``chain.call_batch`` (it is the chain itself in the reference's sampler).
    With ``world > 1`` every rank must call ``map`` with the same ``plist``; each evaluates its shard and
    the log-likelihoods are all-gathered (blobs stay on the owning rank, others get ``None``).
    """

    def __init__(self, chain, rank=0, world=1, device=None):
        self.chain = chain
        self.rank, self.world, self.device = rank, world, device

    def map(self, fn, plist):
        plist = list(plist)
        n = len(plist)
        mine = shard_indices(n, self.rank, self.world)
        res = self.chain.call_batch([plist[i] for i in mine]) if mine else []
        lnl = gather_walker_values(np.array([r[0] for r in res], dtype=np.float64), n, self.rank, self.world,
                                   self.device)
        blobs = {i: r[1] for i, r in zip(mine, res)}
        return [(float(lnl[i]), blobs.get(i)) for i in range(n)]
