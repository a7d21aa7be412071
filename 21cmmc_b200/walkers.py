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


# ------------------------------------------------------------------------------------------------
# Vectorised sampler entry points (the other half): zeus ``vectorize=True`` and UltraNest
# ``vectorized=True`` call the log-probability with ALL points of a step at once (src/py21cmmc/mcmc.py:495-511 builds
# the zeus sampler, :517-545 the UltraNest likelihood / prior).  The reference's vectorised UltraNest branch pushes the
# whole array through one context; here the points become one batched device call.
# ------------------------------------------------------------------------------------------------
def vectorized_log_prob(chain, pool=None, with_blobs=False):
    """``f(P) -> lnL`` for ``P`` of shape (N, ndim) or (ndim,): the ``posterior`` of mcmc.py:473-488, vectorised.

    Points outside the prior box get ``-inf`` without running the cores (mcmc.py:465-470 / ``chain.isValid``).
    ``pool``: a ``BatchedWalkerPool`` to shard the points over ranks; default: this process only.  With
    ``with_blobs`` the return value is ``(lnL, structured blob rows)`` as zeus expects with ``blobs_dtype``.
    """
    mapper = pool.map if pool is not None else (lambda fn, plist: chain.call_batch(plist))

    def log_prob(P):
        P = np.asarray(P, dtype=float)
        single = P.ndim == 1
        res = mapper(None, [list(row) for row in np.atleast_2d(P)])
        lnl = np.array([r[0] for r in res], dtype=np.float64)
        if with_blobs:
            blobs = blobs_to_structured([r[1] for r in res])
            return (lnl[0], blobs[0]) if single else (lnl, blobs)
        return float(lnl[0]) if single else lnl

    return log_prob


def ultranest_prior_transform(params, vectorized=True):
    """Unit cube -> parameter box, mcmc.py:533-543 (``params[i] = [guess, min, max, width]``)."""
    lo = np.array([params[i][1] for i in range(len(params.keys))], dtype=float)
    hi = np.array([params[i][2] for i in range(len(params.keys))], dtype=float)

    def prior(p):
        p = np.asarray(p, dtype=float)
        if vectorized and p.ndim == 1:
            raise ValueError("a vectorised prior transform takes (N, ndim) points")
        return lo + p * (hi - lo)

    return prior


# ------------------------------------------------------------------------------------------------
# HDF5 hand-off: the datasets of the reference's HDFStorage (cosmoHammer.py:60-105 reset,
# :208-246 grow, :248-295 save_step), so that ``py21cmmc.analyse`` / ``HDFStorage.get_blobs`` read what this package's
# samplers wrote.  h5py is imported lazily; without it the structured rows still go to ``BlobChain`` (.npy).
# ------------------------------------------------------------------------------------------------
class HDFBlobStorage:
    """Writer with the group / dataset layout of the reference's ``HDFStorage``.

    Group ``name`` holds ``chain (niter, nwalkers, ndim)``, ``log_prob``, ``trials``, ``trial_log_prob``,
    ``accepted (niter, nwalkers)`` and, once the first blobs arrive, ``blobs (niter, nwalkers)`` with the structured
    dtype ``blob_dtype`` derives; attributes ``iteration``, ``nwalkers``, ``ndim``, ``has_blobs``.
    """

    def __init__(self, filename, name="sample_0"):
        try:
            import h5py  # noqa: F401
        except ImportError as e:  # pragma: no cover - h5py is absent from the build image
            raise ImportError("HDFBlobStorage needs h5py; use BlobChain (.npy) without it") from e
        self.filename, self.name = filename, name

    def open(self, mode="r"):
        import h5py

        return h5py.File(self.filename, mode)

    @staticmethod
    def layout(nwalkers, ndim, blobs=None):
        """{dataset: (initial shape, maxshape, dtype)} -- what ``reset`` + the first ``grow`` create."""
        out = {
            "chain": ((0, nwalkers, ndim), (None, nwalkers, ndim), np.dtype(np.float64)),
            "trials": ((0, nwalkers, ndim), (None, nwalkers, ndim), np.dtype(np.float64)),
            "accepted": ((0, nwalkers), (None, nwalkers), np.dtype(int)),
            "log_prob": ((0, nwalkers), (None, nwalkers), np.dtype(np.float64)),
            "trial_log_prob": ((0, nwalkers), (None, nwalkers), np.dtype(np.float64)),
        }
        if blobs:
            out["blobs"] = ((0, nwalkers), (None, nwalkers), blob_dtype(blobs))
        return out

    def reset(self, nwalkers, params):
        """cosmoHammer.py:60-105; ``params`` = the chain's ``Params`` (names + guesses are stored as ``guess``)."""
        ndim = len(params.keys)
        with self.open("a") as f:
            if self.name in f:
                del f[self.name]
            g = f.create_group(self.name)
            g.attrs["nwalkers"], g.attrs["ndim"] = nwalkers, ndim
            g.attrs["has_blobs"], g.attrs["iteration"] = False, 0
            guess = np.zeros(1, dtype=[(k, np.float64) for k in params.keys])
            for i, k in enumerate(params.keys):
                guess[k] = params[i][0]
            g.create_dataset("guess", data=guess)
            for dsname, (shape, maxshape, dtype) in self.layout(nwalkers, ndim).items():
                g.create_dataset(dsname, shape, maxshape=maxshape, dtype=dtype)

    def grow(self, ngrow, blobs):
        with self.open("a") as f:
            g = f[self.name]
            if g.attrs["has_blobs"] and not blobs:
                raise ValueError("inconsistent use of blobs")
            if g.attrs["iteration"] > 0 and blobs and not g.attrs["has_blobs"]:
                raise ValueError("inconsistent use of blobs")
            ntot = g.attrs["iteration"] + ngrow
            for dsname in ("chain", "trials", "accepted", "log_prob", "trial_log_prob"):
                g[dsname].resize(ntot, axis=0)
            if blobs:
                if not g.attrs["has_blobs"]:
                    nwalkers = g.attrs["nwalkers"]
                    g.create_dataset("blobs", (ntot, nwalkers), maxshape=(None, nwalkers), dtype=blob_dtype(blobs))
                else:
                    g["blobs"].resize(ntot, axis=0)
                g.attrs["has_blobs"] = True

    def save_step(self, coords, log_prob, blobs, truepos, trueprob, accepted, random_state=()):
        with self.open("a") as f:
            g = f[self.name]
            it = g.attrs["iteration"]
            g["chain"][it, :, :] = coords
            g["log_prob"][it, :] = log_prob
            g["trials"][it, :, :] = truepos
            g["trial_log_prob"][it, :] = trueprob
            if blobs is not None and len(blobs) and blobs[0]:
                g["blobs"][it, ...] = blobs_to_structured(blobs, g["blobs"].dtype)
            g["accepted"][it, :] = accepted
            for i, v in enumerate(random_state):
                g.attrs[f"random_state_{i}"] = v
            g.attrs["iteration"] = it + 1
