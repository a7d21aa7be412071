"""Python handle of a ps21 plan: geometry + bins of one box layout, bound to one CUDA device.

Plans are cached per process (``get_plan``) so that likelihood objects stay picklable / YAML-able:
they never hold a device handle (the reference pickles likelihoods into pool workers,
src/py21cmmc/mcmc.py:580-583).
"""
from __future__ import annotations

import collections
import ctypes as C
import os
import threading

import numpy as np

from . import _capi
from . import binning


def _dptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class PowerPlan:
    """Batched power-spectrum / likelihood plan for boxes of shape (n0, n1, n2_total).

    Parameters
    ----------
    shape : (n0, n1, n2_total)
    chunk_bounds : list of (start, end) along the last axis; one spectrum per chunk
    lengths : list of (L0, L1, L2_c) per chunk, Mpc
    nbins, log_bins, ignore_* : as ``compute_power`` (likelihood.py:734-743)
    edge_rule : top-edge rule of the installed powerbox generation ("legacy_1.001" | "inscribed")
    mode : "spherical" | "cylindrical";  path : "auto" | "general" | "fused" | "twopass"
    """

    def __init__(self, shape, chunk_bounds, lengths, nbins=None, log_bins=True, ignore_kperp_zero=True,
                 ignore_kpar_zero=False, ignore_k_zero=False, edge_rule="legacy_1.001", mode="spherical",
                 device=0, path="auto", workspace_bytes=0, library=None):
        self._lib = _capi.load(library)
        n0, n1, n2_total = (int(s) for s in shape)
        self.shape = (n0, n1, n2_total)
        self.chunk_bounds = [(int(a), int(b)) for a, b in chunk_bounds]
        self.nchunks = len(self.chunk_bounds)
        self.lengths = [tuple(float(x) for x in L) for L in lengths]
        if len(self.lengths) != self.nchunks:
            raise ValueError("one (L0, L1, L2) triple per chunk is required")
        self.log_bins = bool(log_bins)
        self.mode = mode
        self.device = int(device)
        cyl = mode == "cylindrical"
        if mode not in ("spherical", "cylindrical"):
            raise ValueError("mode must be 'spherical' or 'cylindrical'")
        nz0 = self.chunk_bounds[0][1] - self.chunk_bounds[0][0]
        if nbins is None:
            nbins = binning.default_nbins((n0, n1, nz0), 2 if cyl else 3)
        self.nbins = int(nbins)

        L0, L1 = self.lengths[0][0], self.lengths[0][1]
        if any(L[0] != L0 or L[1] != L1 for L in self.lengths):
            raise ValueError("all chunks must share the sky-plane lengths")
        k0sq = axis_sq = binning.axis_wavenumbers(n0, L0) ** 2
        k1sq = binning.axis_wavenumbers(n1, L1) ** 2
        k2sq, edges, centres, volume = [], [], [], []
        for (a, b), L in zip(self.chunk_bounds, self.lengths):
            nz = b - a
            kz2 = binning.axis_wavenumbers(nz, L[2]) ** 2
            k2sq.append(kz2)
            axes = [k0sq, k1sq] if cyl else [k0sq, k1sq, kz2]
            e = binning.bin_edges(axes, self.nbins, self.log_bins, edge_rule)
            edges.append(e)
            centres.append(binning.bin_centres(e, self.log_bins))
            volume.append(float(np.prod(np.array(L, dtype=np.float64))))
        del axis_sq
        self.edges = np.ascontiguousarray(np.stack(edges), dtype=np.float64)
        self.kcentres = np.ascontiguousarray(np.stack(centres), dtype=np.float64)
        self.kpar = [np.sqrt(k[: (len(k) // 2 + 1)]) for k in k2sq]  # |k_par| of the half spectrum
        self._keep = dict(
            cs=np.ascontiguousarray([a for a, _ in self.chunk_bounds], dtype=np.int32),
            ce=np.ascontiguousarray([b for _, b in self.chunk_bounds], dtype=np.int32),
            vol=np.ascontiguousarray(volume, dtype=np.float64),
            k0=np.ascontiguousarray(k0sq, dtype=np.float64),
            k1=np.ascontiguousarray(k1sq, dtype=np.float64),
            k2=np.ascontiguousarray(np.concatenate(k2sq), dtype=np.float64),
        )
        d = _capi.Desc()
        d.n0, d.n1, d.n2_total, d.nchunks = n0, n1, n2_total, self.nchunks
        i32p = C.POINTER(C.c_int32)
        d.chunk_start = self._keep["cs"].ctypes.data_as(i32p)
        d.chunk_end = self._keep["ce"].ctypes.data_as(i32p)
        d.volume = _dptr(self._keep["vol"])
        d.k0sq, d.k1sq, d.k2sq = _dptr(self._keep["k0"]), _dptr(self._keep["k1"]), _dptr(self._keep["k2"])
        d.nbins = self.nbins
        d.edges, d.kcentres = _dptr(self.edges), _dptr(self.kcentres)
        d.ignore_kperp_zero, d.ignore_kpar_zero, d.ignore_k_zero = (
            int(bool(ignore_kperp_zero)), int(bool(ignore_kpar_zero)), int(bool(ignore_k_zero)))
        d.mode = _capi.PS21_MODE_CYLINDRICAL if cyl else _capi.PS21_MODE_SPHERICAL
        d.device = self.device
        d.path = {"auto": _capi.PS21_PATH_AUTO, "general": _capi.PS21_PATH_GENERAL,
                  "fused": _capi.PS21_PATH_FUSED, "twopass": _capi.PS21_PATH_TWOPASS}[path]
        d.workspace_bytes = int(workspace_bytes)
        h = C.c_void_p()
        _capi.check(self._lib.ps21_plan_create(C.byref(d), C.byref(h)))
        self._h = h
        offs = (C.c_int64 * (self.nchunks + 1))()
        _capi.check(self._lib.ps21_plan_layout(self._h, offs))
        self.offsets = [int(o) for o in offs]
        self.out_per_box = self.offsets[-1]
        cnt = np.zeros(self.out_per_box, dtype=np.int64)
        _capi.check(self._lib.ps21_plan_counts(self._h, cnt.ctypes.data_as(C.POINTER(C.c_int64))))
        self.counts = cnt
        self._has_data = False
        self._data_key = None
        # re-entrant: a likelihood holds it across set_data + lnl so that two likelihoods sharing this (cached)
        # plan cannot evaluate against each other's data tables
        self._lock = threading.RLock()

    # ------------------------------------------------------------------ bookkeeping
    def close(self):
        if getattr(self, "_h", None):
            self._lib.ps21_plan_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def info(self):
        inf = _capi.Info()
        _capi.check(self._lib.ps21_plan_info(self._h, C.byref(inf)))
        return {f[0]: getattr(inf, f[0]) for f in _capi.Info._fields_}

    KERNEL_KINDS = ("offset_estimate", "fft_pass1", "fft_pass_mid", "fft_last_pass_bin", "finalize")

    def set_profiling(self, on=True):
        _capi.check(self._lib.ps21_plan_set_profiling(self._h, int(bool(on))))

    def set_graphs(self, on=True):
        """Replay the kernel sequence of repeated ``*_device`` calls as one CUDA graph (ps21_plan_set_graphs)."""
        _capi.check(self._lib.ps21_plan_set_graphs(self._h, int(bool(on))))

    def graph_stats(self):
        cap, rep = C.c_int64(), C.c_int64()
        _capi.check(self._lib.ps21_plan_graph_stats(self._h, C.byref(cap), C.byref(rep)))
        return {"captures": cap.value, "replays": rep.value}

    def release_workspace(self):
        """Free scratch / staging / result buffers (they regrow on demand); the plan stays usable."""
        with self._lock:
            if getattr(self, "_h", None):
                _capi.check(self._lib.ps21_plan_release_workspace(self._h))

    def profile(self):
        """{kind: (total ms, launches)} since the last query (CUDA events on the launching stream)."""
        ms = (C.c_double * 5)()
        n = (C.c_int64 * 5)()
        _capi.check(self._lib.ps21_plan_profile(self._h, ms, n))
        return {k: (float(ms[i]), int(n[i])) for i, k in enumerate(self.KERNEL_KINDS)}

    @property
    def path(self):
        return {1: "general", 2: "fused", 3: "twopass"}[self.info()["path_of_chunk0"]]

    def chunk_counts(self, c):
        return self._chunk_view(self.counts, c)

    def _chunk_view(self, rec, c):
        """Slice one chunk out of a (..., out_per_box) record, shaped (nbins,) or (nbins, nh)."""
        a, b = self.offsets[c], self.offsets[c + 1]
        v = rec[..., a:b]
        if self.mode == "cylindrical":
            return v.reshape(v.shape[:-1] + (self.nbins, (b - a) // self.nbins))
        return v

    def box_elems(self):
        return self.shape[0] * self.shape[1] * self.shape[2]

    # ------------------------------------------------------------------ host-buffer API
    def _as_host_batch(self, boxes):
        a = np.asarray(boxes)
        if a.dtype != np.float32:
            a = a.astype(np.float32)
        if a.ndim == 3:
            a = a[None]
        if a.shape[1:] != self.shape:
            raise ValueError(f"boxes of shape {a.shape[1:]} do not match the plan {self.shape}")
        return np.ascontiguousarray(a)

    def _box_pointers(self, boxes):
        """(pointer array, keep-alive list) for a sequence of per-walker host boxes; no gather into one buffer.

        A box that already is a C-contiguous float32 (n0, n1, n2_total) array is passed as it is (zero copy); anything
        else (another dtype, a strided view) is converted first, which is the only copy made on the host."""
        keep = []
        for b in boxes:
            a = np.asarray(b)
            if a.shape != self.shape:
                raise ValueError(f"box of shape {a.shape} does not match the plan {self.shape}")
            if a.dtype != np.float32 or not a.flags.c_contiguous:
                a = np.ascontiguousarray(a, dtype=np.float32)
            keep.append(a)
        ptrs = (C.c_void_p * len(keep))(*[a.ctypes.data for a in keep])
        return ptrs, keep

    def power_host(self, boxes):
        """Power of every chunk of every box -> (batch, out_per_box).

        boxes: (batch, n0, n1, n2_total) array, one box, or a list / tuple of per-walker boxes (passed by pointer)."""
        if isinstance(boxes, (list, tuple)):
            ptrs, keep = self._box_pointers(boxes)
            out = np.empty((len(keep), self.out_per_box), dtype=np.float64)
            with self._lock:
                _capi.check(self._lib.ps21_power_host_ptrs(self._h, ptrs, len(keep), out.ctypes.data))
            return out
        a = self._as_host_batch(boxes)
        out = np.empty((a.shape[0], self.out_per_box), dtype=np.float64)
        with self._lock:
            _capi.check(self._lib.ps21_power_host(self._h, a.ctypes.data, a.shape[0], self.box_elems(),
                                                  out.ctypes.data))
        return out

    def set_data(self, data, noise=None, kmask=None, model_uncertainty=0.15, error_on_model=True):
        """data/noise/kmask: (nsets, nchunks, nbins) arrays evaluated at this plan's bin centres."""
        data = np.ascontiguousarray(data, dtype=np.float64).reshape(-1, self.out_per_box)
        nsets = data.shape[0]
        noise_p = None
        if noise is not None:
            noise = np.ascontiguousarray(noise, dtype=np.float64).reshape(nsets, self.out_per_box)
            noise_p = _dptr(noise)
        if kmask is None:
            kmask = np.ones_like(data, dtype=np.uint8)
        kmask = np.ascontiguousarray(kmask, dtype=np.uint8).reshape(nsets, self.out_per_box)
        # the tables are walker independent: skip the upload when they are the ones already on the device
        key = (nsets, data.tobytes(), None if noise is None else noise.tobytes(), kmask.tobytes(),
               float(model_uncertainty), bool(error_on_model))
        with self._lock:
            if key == self._data_key:
                return
            _capi.check(self._lib.ps21_set_data(self._h, nsets, _dptr(data), noise_p,
                                                kmask.ctypes.data_as(C.POINTER(C.c_uint8)),
                                                float(model_uncertainty), int(bool(error_on_model))))
            self._data_key = key
            self._has_data = True

    def lnl_host(self, boxes, want_delta=False, want_power=False):
        """(lnL[batch], Delta^2 or None, power or None) for host boxes (array, or list of per-walker arrays)."""
        by_ptr = isinstance(boxes, (list, tuple))
        if by_ptr:
            ptrs, keep = self._box_pointers(boxes)
            n = len(keep)
        else:
            a = self._as_host_batch(boxes)
            n = a.shape[0]
        lnl = np.empty(n, dtype=np.float64)
        delta = np.empty((n, self.out_per_box), dtype=np.float64) if want_delta else None
        power = np.empty((n, self.out_per_box), dtype=np.float64) if want_power else None
        dp = delta.ctypes.data if want_delta else None
        pp = power.ctypes.data if want_power else None
        with self._lock:
            if by_ptr:
                _capi.check(self._lib.ps21_lnl_host_ptrs(self._h, ptrs, n, lnl.ctypes.data, dp, pp))
            else:
                _capi.check(self._lib.ps21_lnl_host(self._h, a.ctypes.data, n, self.box_elems(), lnl.ctypes.data,
                                                    dp, pp))
        return lnl, delta, power

    def lnl_from_delta(self, delta):
        d = np.ascontiguousarray(delta, dtype=np.float64).reshape(-1, self.out_per_box)
        out = np.empty(d.shape[0], dtype=np.float64)
        with self._lock:
            _capi.check(self._lib.ps21_lnl_from_delta(self._h, _dptr(d), d.shape[0], _dptr(out)))
        return out

    # ------------------------------------------------------------------ device-buffer API (torch = allocator)
    def _check_dev(self, t):
        import torch

        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise TypeError("boxes must be a contiguous float32 CUDA tensor")
        if t.dim() == 3:
            t = t.unsqueeze(0)
        if tuple(t.shape[1:]) != self.shape:
            raise ValueError(f"boxes of shape {tuple(t.shape[1:])} do not match the plan {self.shape}")
        if t.device.index != self.device:
            raise ValueError("tensor lives on another device than the plan")
        return t

    def power_device(self, boxes, out=None, stream=None):
        import torch

        t = self._check_dev(boxes)
        if out is None:
            out = torch.empty((t.shape[0], self.out_per_box), dtype=torch.float64, device=t.device)
        s = torch.cuda.current_stream(t.device).cuda_stream if stream is None else stream
        with self._lock:
            _capi.check(self._lib.ps21_power_batched(self._h, t.data_ptr(), t.shape[0], self.box_elems(),
                                                     out.data_ptr(), s))
        return out

    def lnl_device(self, boxes, lnl=None, delta=None, power=None, stream=None):
        import torch

        t = self._check_dev(boxes)
        if lnl is None:
            lnl = torch.empty((t.shape[0],), dtype=torch.float64, device=t.device)
        s = torch.cuda.current_stream(t.device).cuda_stream if stream is None else stream
        with self._lock:
            _capi.check(self._lib.ps21_lnl_batched(
                self._h, t.data_ptr(), t.shape[0], self.box_elems(), lnl.data_ptr(),
                delta.data_ptr() if delta is not None else None,
                power.data_ptr() if power is not None else None, s))
        return lnl


# Per-process plan cache, least recently used first.  A plan keeps its bin-index tables and a half-spectrum scratch
# that grows up to 1/6 of the device memory, so the cache is bounded (PS21_PLAN_CACHE, default 8 plans): an evicted
# plan gives its workspace back at once and is destroyed when the last likelihood using it lets go of it.
_PLANS = collections.OrderedDict()
_PLANS_LOCK = threading.Lock()
_PLAN_CACHE_MAX = max(1, int(os.environ.get("PS21_PLAN_CACHE", "8")))


def get_plan(shape, chunk_bounds, lengths, nbins=None, log_bins=True, ignore_kperp_zero=True,
             ignore_kpar_zero=False, ignore_k_zero=False, edge_rule="legacy_1.001", mode="spherical",
             device=0, path="auto", library=None) -> PowerPlan:
    """Per-process plan cache keyed by the complete geometry."""
    library = library or _capi.LIB_PATH
    key = (tuple(shape), tuple(map(tuple, chunk_bounds)), tuple(map(tuple, lengths)), nbins, bool(log_bins),
           bool(ignore_kperp_zero), bool(ignore_kpar_zero), bool(ignore_k_zero), edge_rule, mode, int(device),
           path, library)
    with _PLANS_LOCK:
        pl = _PLANS.get(key)
        if pl is not None:
            _PLANS.move_to_end(key)
            return pl
        pl = PowerPlan(shape, chunk_bounds, lengths, nbins, log_bins, ignore_kperp_zero, ignore_kpar_zero,
                       ignore_k_zero, edge_rule, mode, device, path, library=library)
        _PLANS[key] = pl
        while len(_PLANS) > _PLAN_CACHE_MAX:
            _, old = _PLANS.popitem(last=False)
            try:
                old.release_workspace()
            except Exception:
                pass
    return pl


def clear_plans():
    with _PLANS_LOCK:
        for pl in _PLANS.values():
            pl.close()
        _PLANS.clear()


# ---------------------------------------------------------------------------------------------------------
# page-locked host buffers for walker lightcones (copies from them run at the PCIe line rate)
# ---------------------------------------------------------------------------------------------------------
class _PinnedOwner:
    def __init__(self, lib, ptr):
        self.lib, self.ptr = lib, ptr

    def __del__(self):
        try:
            self.lib.ps21_host_free(self.ptr)
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float32, library=None):
    """Uninitialised NumPy array in page-locked memory (cudaHostAlloc through the C ABI); freed with the array."""
    lib = _capi.load(library)
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    ptr = C.c_void_p()
    _capi.check(lib.ps21_host_alloc(C.byref(ptr), max(nbytes, 1)))
    owner = _PinnedOwner(lib, ptr)
    buf = (C.c_char * max(nbytes, 1)).from_address(ptr.value)
    buf._owner = owner   # the ctypes buffer keeps the allocation alive, the array keeps the buffer alive
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


def pin_array(a, library=None):
    """Page-lock an existing C-contiguous array in place (cudaHostRegister); returns an unpin callable."""
    lib = _capi.load(library)
    a = np.asarray(a)
    if not a.flags.c_contiguous:
        raise ValueError("only C-contiguous arrays can be pinned")
    _capi.check(lib.ps21_host_register(a.ctypes.data, a.nbytes))
    return lambda: _capi.check(lib.ps21_host_unregister(a.ctypes.data))
