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


# ------------------------------------------------------------------------------------------------
# Blob hand-off: the per-walker ``k_i`` / ``delta_i`` records that ``store`` puts
# into the context's data dict are persisted by the reference's HDFStorage as ONE structured array per
# iteration (src/py21cmmc/cosmoHammer.py:227-242 builds the dtype, :281-288 packs a step).
# ------------------------------------------------------------------------------------------------
def blob_dtype(blobs: dict):
    """Structured dtype of a blob dict, exactly as ``HDFStorage.grow`` derives it (cosmoHammer.py:231-236)."""
    dt = []
    for k, v in blobs.items():
        a = np.atleast_1d(v)
        shape = a.shape
        if len(shape) == 1:
            shape = shape[0]
        dt.append((k, (a.dtype, shape)))
    return np.dtype(dt)


def blobs_to_structured(blob_dicts, dtype=None):
    """One structured row per walker, as ``HDFStorage.save_step`` packs a step (cosmoHammer.py:281-285).

    ``blob_dicts``: the second element of every ``(lnL, blobs)`` result of ``chain.call_batch`` /
    ``BatchedWalkerPool.map``; rejected walkers (``[]`` / ``None``) become NaN rows.
    """
    first = next((b for b in blob_dicts if b), None)
    if first is None:
        return None
    dtype = dtype or blob_dtype(first)
    out = np.zeros(len(blob_dicts), dtype=dtype)
    for name in dtype.names:
        if np.issubdtype(dtype[name].base, np.floating):
            out[name] = np.nan
    for i, b in enumerate(blob_dicts):
        if b:
            out[i] = tuple(b[name] for name in dtype.names)
    return out


class BlobChain:
    """Grow-by-iteration store of structured blob rows: (niter, nwalkers), the layout of the reference's
    ``blobs`` dataset.  ``save(path)`` writes a ``.npy`` that ``np.load`` returns with named fields, so
    ``blobs["delta_0"][it, walker]`` reads like the reference's ``samples.get_blobs()``."""

    def __init__(self):
        self.rows = []
        self.dtype = None

    def save_step(self, blob_dicts):
        rec = blobs_to_structured(blob_dicts, self.dtype)
        if rec is None:
            if self.rows:
                raise ValueError("inconsistent use of blobs")  # cosmoHammer.py:297-301
            return
        if self.dtype is None:
            self.dtype = rec.dtype
        if self.rows and len(rec) != len(self.rows[0]):
            raise ValueError("number of walkers changed between iterations")
        self.rows.append(rec)

    def array(self):
        return np.stack(self.rows) if self.rows else None

    def save(self, path):
        np.save(path, self.array())
