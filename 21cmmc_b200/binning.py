"""Host-side geometry of the power-spectrum estimator: wavenumbers, bin edges, bin centres.

Everything that decides a bin index is produced here with the NumPy expressions powerbox itself
evaluates (``powerbox.dft.fftfreq``, ``powerbox.tools._magnitude_grid`` / ``_get_binweights`` /
``_getbins`` [third party, restated in SURVEY.md Appendix A]) and handed to the device as explicit
fp64 arrays, so bin assignment is bit-exact.  Only 1-D arrays are needed: floating-point addition
and ``sqrt`` are monotone, hence the extrema of ``sqrt((kx^2+ky^2)+kz^2)`` over the grid are reached
at the per-axis extrema.
"""
from __future__ import annotations

import numpy as np

EDGE_RULES = ("legacy_1.001", "inscribed")


def chunk_indices(n_slices: int, nchunks: int):
    """Chunk boundaries of ``Likelihood1DPowerLightcone.reduce_data`` (likelihood.py:836-847)."""
    idx = list(range(0, n_slices, round(n_slices / nchunks)))
    if len(idx) > nchunks:
        idx = idx[:-1]
    idx.append(n_slices)
    return idx


def axis_wavenumbers(n: int, length: float) -> np.ndarray:
    """k of one axis in *unshifted* FFT order, rad/Mpc (powerbox.dft.fftfreq with b=1, un-shifted)."""
    d = float(length) / float(n)
    return np.fft.fftfreq(n, d=d) * (2 * np.pi / 1.0)


def default_nbins(shape, res_ndim=None) -> int:
    """powerbox.get_power's default: ``int(prod(N[:res_ndim]) ** (1/res_ndim) / 2.2)``."""
    res_ndim = len(shape) if res_ndim is None else res_ndim
    return int(np.prod(shape[:res_ndim]) ** (1.0 / res_ndim) / 2.2)


def bin_edges(ksq_axes, nbins: int, log_bins: bool, edge_rule: str = "legacy_1.001") -> np.ndarray:
    """Edges of the |k| bins for the axes whose squared wavenumbers are given.

    legacy_1.001: ``linspace/logspace(min, 1.001*max|k|)`` (powerbox <= 0.6, the form in use when the
    reference pinned ``powerbox>=0.5.7``); inscribed: top edge = smallest per-axis Nyquist
    (powerbox >= 0.7 ``_getbins``).
    """
    if edge_rule not in EDGE_RULES:
        raise ValueError(f"edge_rule must be one of {EDGE_RULES}")
    maxs = [np.max(k2) for k2 in ksq_axes]
    zero = np.float64(0.0)
    if edge_rule == "legacy_1.001":
        acc = maxs[0]
        for m in maxs[1:]:
            acc = acc + m  # ((a+b)+c): the association np.sum(axis=0) uses
        top = np.sqrt(acc) * 1.001
    else:
        tops = []
        for i, m in enumerate(maxs):
            acc = m if i == 0 else zero
            for j in range(1, len(maxs)):
                acc = acc + (m if j == i else zero)
            tops.append(np.sqrt(acc))
        top = np.min(tops)
    if not log_bins:
        return np.linspace(zero, top, nbins + 1)
    mins = []
    for k2 in ksq_axes:
        pos = k2[k2 > 0]
        if pos.size:
            mins.append(np.sqrt(pos.min()))
    mn = np.min(mins)
    return np.logspace(np.log10(mn), np.log10(top), nbins + 1)


def detect_edge_rule():
    """Edge rule of the powerbox installed next to this package, or None when powerbox is not importable.

    SURVEY.md section 8c: the top-edge rule changed between powerbox generations, so it is probed rather than
    assumed: a small anisotropic white-noise box is run through ``powerbox.get_power`` and the returned edges are
    compared with both rules.
    """
    try:
        from powerbox.tools import get_power
    except Exception:
        return None
    shape, length = (8, 8, 12), (8.0, 8.0, 6.0)
    box = np.random.default_rng(0).normal(size=shape).astype(np.float32)
    edges = np.asarray(get_power(box, boxlength=list(length), bins=5, bin_ave=False, get_variance=False,
                                 log_bins=False)[1])
    k2 = [axis_wavenumbers(n, l) ** 2 for n, l in zip(shape, length)]
    for rule in EDGE_RULES:
        if np.allclose(edges, bin_edges(k2, 5, False, rule), rtol=1e-12, atol=0.0):
            return rule
    raise RuntimeError("installed powerbox matches neither known bin-edge rule; pass explicit edges")


def bin_centres(edges: np.ndarray, log_bins: bool) -> np.ndarray:
    """likelihood.py:792-797."""
    k = np.asarray(edges)
    if log_bins:
        return np.exp((np.log(k[1:]) + np.log(k[:-1])) / 2)
    return (k[1:] + k[:-1]) / 2
