"""CPU oracle for the 21CMMC power-spectrum-likelihood hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the arithmetic of this path lives in the third-party package ``powerbox``
(``powerbox.tools.get_power``; the reference pins it only as ``powerbox>=0.5.7``,
/root/reference/pyproject.toml:34), which is absent from /root/reference and cannot be installed
here.  The reference's own tests hold no golden vector / known-answer value for the path
(SURVEY.md section 8c).  This file therefore *restates* powerbox's published algorithm
(``dft.fft`` -> ``|FT|^2/V`` -> ``angular_average_nd`` with ``np.digitize``/``np.bincount``) and is
pinned by analytic known-answer tests (white noise, plane waves, Parseval) in
tests/test_oracle.py, plus the frozen vectors in tests/golden/.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module.  Nothing under ``21cmmc_b200/`` (the product) imports it.

Every function cites the reference call site (file:line relative to /root/reference) or the
powerbox routine ("[3P]") that it follows.
"""
from __future__ import annotations

import numpy as np

np.seterr(invalid="ignore", divide="ignore")  # src/py21cmmc/likelihood.py:22

EDGE_RULES = ("legacy_1.001", "inscribed")


# --------------------------------------------------------------------------------------
# powerbox.dft [3P]
# --------------------------------------------------------------------------------------
def fftfreq(n: int, d: float) -> np.ndarray:
    """powerbox.dft.fftfreq with b=1 [3P]: ``fftshift(np.fft.fftfreq(N, d)) * 2*pi`` (rad/Mpc)."""
    return np.fft.fftshift(np.fft.fftfreq(n, d=d)) * (2 * np.pi / 1.0)


def dft_power(box: np.ndarray, boxlength, fft_dtype=np.float64):
    """powerbox.dft.fft(a=1, b=1) followed by get_power's ``P = real(FT conj(FT))/V**2 * V`` [3P].

    Returns the *shifted* full-grid power P[k] = (V/N^2) |sum_x box(x) exp(-ik.x)|^2 and the
    shifted per-axis wavenumbers.  ``fft_dtype=np.float32`` reproduces what NumPy>=2 / pyFFTW do
    for float32 input (complex64 transform); float64 is the truth used for parity tolerances.
    """
    box = np.asarray(box)
    shape = box.shape
    L = np.array([float(boxlength)] * box.ndim if np.isscalar(boxlength) else [float(x) for x in boxlength])
    V = float(np.prod(L))
    Vx = V / np.prod(shape)
    ft = Vx * np.fft.fftshift(np.fft.fftn(box.astype(fft_dtype)))
    dx = np.array([float(l) / float(n) for l, n in zip(L, shape)])
    freq = [fftfreq(n, d=d) for n, d in zip(shape, dx)]
    # powerbox's left-edge phase factors have unit modulus and cancel in |FT|^2.
    P = np.real(ft * np.conj(ft) / V**2).astype(np.float64)
    P *= V
    return P, freq


# --------------------------------------------------------------------------------------
# powerbox.tools [3P]
# --------------------------------------------------------------------------------------
def magnitude_grid(coords) -> np.ndarray:
    """powerbox.tools._magnitude_grid [3P]: sqrt(sum of squares) with sequential axis-0 summation."""
    sq = np.meshgrid(*[np.asarray(c) ** 2 for c in coords], indexing="ij")
    return np.sqrt(np.sum(sq, axis=0))


def get_bins(nbins: int, coords: np.ndarray, log_bins: bool, edge_rule: str = "legacy_1.001") -> np.ndarray:
    """Bin edges.  legacy: powerbox.tools._get_binweights (<=0.6.x); inscribed: ``_getbins`` (>=0.7) [3P]."""
    if edge_rule == "legacy_1.001":
        top = coords.max() * 1.001
    elif edge_rule == "inscribed":
        maxs = [np.max(coords, axis=i) for i in range(coords.ndim)]
        flat = []
        for m in maxs:
            flat.extend(np.ravel(m))
        top = np.min(flat)
    else:
        raise ValueError(f"edge_rule must be one of {EDGE_RULES}")
    if not log_bins:
        return np.linspace(coords.min(), top, nbins + 1)
    mn = coords[coords > 0].min()
    return np.logspace(np.log10(mn), np.log10(top), nbins + 1)


def angular_average(field, coords_grid, bins, weights, log_bins=False, edge_rule="legacy_1.001"):
    """powerbox.tools.angular_average with bin_ave=False, average=True [3P].

    Returns (res, edges, sumweights).  Empty bins give 0/0 = NaN, as in the reference.
    """
    edges = get_bins(bins, coords_grid, log_bins, edge_rule) if not np.iterable(bins) else np.asarray(bins)
    indx = np.digitize(coords_grid.ravel(), edges)
    nb = len(edges) - 1
    sumweights = np.bincount(indx, weights=np.asarray(weights, dtype=np.float64).ravel(), minlength=nb + 2)[1:-1]
    fw = (field * weights).ravel()
    res = np.bincount(indx, weights=fw, minlength=nb + 2)[1:-1] / sumweights
    return res, edges, sumweights


def get_power(deltax, boxlength, bins=None, log_bins=False, k_weights=1, res_ndim=None,
              edge_rule="legacy_1.001", fft_dtype=np.float64, return_counts=False):
    """powerbox.tools.get_power(bin_ave=False, get_variance=False) for gridded input [3P].

    Call sites: src/py21cmmc/likelihood.py:480-488 and :782-790.  No mean subtraction and no
    division by the mean happen for gridded input.  Returns (P[nb], edges[nb+1]) for the
    spherical average, or (P[nb, N2], kperp_edges, [kpar]) when ``res_ndim=2`` (cylindrical).
    """
    deltax = np.asarray(deltax)
    dim = deltax.ndim
    P, freq = dft_power(deltax, boxlength, fft_dtype=fft_dtype)
    if res_ndim is None:
        res_ndim = dim
    if bins is None:
        bins = int(np.prod(deltax.shape[:res_ndim]) ** (1.0 / res_ndim) / 2.2)
    kmag = magnitude_grid(freq[:res_ndim])
    if np.isscalar(k_weights):
        k_weights = np.ones_like(kmag)
    if res_ndim == dim:
        res, edges, sw = angular_average(P, kmag, bins, k_weights, log_bins, edge_rule)
        return (res, edges, sw) if return_counts else (res, edges)
    # angular_average_nd: bin the first res_ndim axes, independently for every remaining column.
    edges = get_bins(bins, kmag, log_bins, edge_rule) if not np.iterable(bins) else np.asarray(bins)
    n1 = int(np.prod(deltax.shape[:res_ndim]))
    n2 = int(np.prod(deltax.shape[res_ndim:]))
    Pf = P.reshape(n1, n2)
    out = np.zeros((len(edges) - 1, n2))
    sws = np.zeros((len(edges) - 1, n2))
    for i in range(n2):
        out[:, i], _, sws[:, i] = angular_average(Pf[:, i].reshape(kmag.shape), kmag, edges, k_weights)
    out = out.reshape((len(edges) - 1,) + deltax.shape[res_ndim:])
    if return_counts:
        return out, edges, freq[res_ndim:], sws.reshape(out.shape)
    return out, edges, freq[res_ndim:]


# --------------------------------------------------------------------------------------
# py21cmmc.likelihood
# --------------------------------------------------------------------------------------
def k_weights_lightcone(shape, ignore_kperp_zero=True, ignore_kpar_zero=False, ignore_k_zero=False):
    """src/py21cmmc/likelihood.py:771-780 (n0 indexes *both* perpendicular axes, as in the reference)."""
    w = np.ones(shape, dtype=int)
    n0 = shape[0]
    n1 = shape[-1]
    if ignore_kperp_zero:
        w[n0 // 2, n0 // 2, :] = 0
    if ignore_kpar_zero:
        w[:, :, n1 // 2] = 0
    if ignore_k_zero:
        w[n0 // 2, n0 // 2, n1 // 2] = 0
    return w


def k_weights_coeval(shape, ignore_kperp_zero=True, ignore_kpar_zero=False, ignore_k_zero=False):
    """src/py21cmmc/likelihood.py:470-478 (n = shape[0] used for all three axes)."""
    w = np.ones(shape, dtype=int)
    n = shape[0]
    if ignore_kperp_zero:
        w[n // 2, n // 2, :] = 0
    if ignore_kpar_zero:
        w[:, :, n // 2] = 0
    if ignore_k_zero:
        w[n // 2, n // 2, n // 2] = 0
    return w


def edges_to_centres(edges, log_bins):
    """src/py21cmmc/likelihood.py:792-797."""
    k = np.asarray(edges)
    if log_bins:
        return np.exp((np.log(k[1:]) + np.log(k[:-1])) / 2)
    return (k[1:] + k[:-1]) / 2


def compute_power_lightcone(box, length, n_psbins, log_bins=True, ignore_kperp_zero=True,
                            ignore_kpar_zero=False, ignore_k_zero=False,
                            edge_rule="legacy_1.001", fft_dtype=np.float64, return_counts=False):
    """Likelihood1DPowerLightcone.compute_power, src/py21cmmc/likelihood.py:734-800."""
    w = k_weights_lightcone(box.shape, ignore_kperp_zero, ignore_kpar_zero, ignore_k_zero)
    res = list(get_power(box, boxlength=length, bins=n_psbins, log_bins=log_bins, k_weights=w,
                         edge_rule=edge_rule, fft_dtype=fft_dtype, return_counts=return_counts))
    res[1] = edges_to_centres(res[1], log_bins)
    return res


def compute_power_coeval(box, L, n_psbins, log_bins=True, ignore_kperp_zero=True,
                         ignore_kpar_zero=False, ignore_k_zero=False,
                         edge_rule="legacy_1.001", fft_dtype=np.float64, return_counts=False):
    """Likelihood1DPowerCoeval.compute_power, src/py21cmmc/likelihood.py:433-498."""
    w = k_weights_coeval(box.shape, ignore_kperp_zero, ignore_kpar_zero, ignore_k_zero)
    res = list(get_power(box, boxlength=L, bins=n_psbins, log_bins=log_bins, k_weights=w,
                         edge_rule=edge_rule, fft_dtype=fft_dtype, return_counts=return_counts))
    res[1] = edges_to_centres(res[1], log_bins)
    return res


def chunk_indices(n_slices: int, nchunks: int):
    """src/py21cmmc/likelihood.py:836-847 (Python banker's ``round``; last chunk absorbs the rest)."""
    idx = list(range(0, n_slices, round(n_slices / nchunks)))
    if len(idx) > nchunks:
        idx = idx[:-1]
    idx.append(n_slices)
    return idx


def reduce_data_lightcone(brightness_temp, box_len, cell_size, nchunks=1, n_psbins=None, logk=True,
                          ignore_kperp_zero=True, ignore_kpar_zero=False, ignore_k_zero=False,
                          edge_rule="legacy_1.001", fft_dtype=np.float64):
    """Likelihood1DPowerLightcone.reduce_data, non-EMU branch, src/py21cmmc/likelihood.py:833-865."""
    n_slices = brightness_temp.shape[-1]
    ci = chunk_indices(n_slices, nchunks)
    data = []
    for i in range(nchunks):
        start, end = ci[i], ci[i + 1]  # IndexError when too few starts, exactly like the reference
        chunklen = (end - start) * cell_size
        power, k = compute_power_lightcone(
            brightness_temp[:, :, start:end], (box_len, box_len, chunklen), n_psbins,
            log_bins=logk, ignore_kperp_zero=ignore_kperp_zero, ignore_kpar_zero=ignore_kpar_zero,
            ignore_k_zero=ignore_k_zero, edge_rule=edge_rule, fft_dtype=fft_dtype)
        data.append({"k": k, "delta": power * k**3 / (2 * np.pi**2)})
    return data


def reduce_data_coeval(boxes, box_len, n_psbins=None, logk=True, ignore_kperp_zero=True,
                       ignore_kpar_zero=False, ignore_k_zero=False,
                       edge_rule="legacy_1.001", fft_dtype=np.float64):
    """Likelihood1DPowerCoeval.reduce_data, non-EMU branch, src/py21cmmc/likelihood.py:639-654."""
    data = []
    for bt in boxes:
        power, k = compute_power_coeval(
            bt, box_len, n_psbins, log_bins=logk, ignore_k_zero=ignore_k_zero,
            ignore_kpar_zero=ignore_kpar_zero, ignore_kperp_zero=ignore_kperp_zero,
            edge_rule=edge_rule, fft_dtype=fft_dtype)
        data.append({"k": k, "delta": power * k**3 / (2 * np.pi**2)})
    return data


def compute_likelihood(model, data, noise=None, min_k=0.1, max_k=1.0, model_uncertainty=0.15,
                       error_on_model=True):
    """Likelihood1DPowerCoeval.computeLikelihood, non-EMU branch, src/py21cmmc/likelihood.py:584-607.

    ``data`` / ``noise`` are lists of dicts with keys k, delta / k, errs
    (splines as in src/py21cmmc/likelihood.py:416-431).
    """
    from scipy.interpolate import InterpolatedUnivariateSpline

    data_spline = [InterpolatedUnivariateSpline(d["k"], d["delta"], k=1) for d in data]
    noise_spline = [InterpolatedUnivariateSpline(n["k"], n["errs"], k=1) for n in noise] if noise else None
    lnl = 0
    nz = 0
    for i, (m, pd) in enumerate(zip(model, data_spline)):
        mask = np.logical_and(m["k"] <= max_k, m["k"] >= min_k)
        moduncert = (model_uncertainty * pd(m["k"][mask]) if not error_on_model
                     else model_uncertainty * m["delta"][mask])
        if noise_spline:
            nz = noise_spline[i](m["k"][mask])
        lnl += -0.5 * np.sum((m["delta"][mask] - pd(m["k"][mask])) ** 2 / (moduncert**2 + nz**2))
    return lnl


def upper_limit_likelihood(hera_data, redshifts, delta, delta_err=None, gaussian=False):
    """Likelihood1DPowerLightconeUpper.computeLikelihood, src/py21cmmc/likelihood.py:2249-2335; with
    ``gaussian=True`` the measured-band-power branch of Likelihood1DPowerCoeval.computeLikelihood, :517-583
    (identical window matvec and error model, Gaussian chi^2 with np.sum instead of erf with nansum).

    ``hera_data``: dict with ``band{z}`` (nfields, nk, 3) = [k, limit, variance] (NaN-padded k),
    ``wfband{z}`` (nfields, nk, nkwf), ``kwfband{z}`` (nkwf,).  ``delta``: (N, nbands, nkwf_max) model
    Delta^2 at the window-function k of every band.  Returns lnL of the N models.
    """
    from scipy.special import erf

    redshifts = np.asarray(redshifts)
    N = delta.shape[0]
    lnl = np.zeros(N)
    for i in range(N):
        for band in redshifts:
            key = str(round(band))
            for field in range(hera_data["band" + key].shape[0]):
                ks = hera_data["band" + key][field, :, 0]
                nk = len(ks[~np.isnan(ks)])
                vals = hera_data["band" + key][field, :nk, 1]
                var = hera_data["band" + key][field, :nk, 2]
                nkwf = len(hera_data["kwfband" + key])
                wf = hera_data["wfband" + key][field, :nk, :].reshape([nk, nkwf])
                zbin = np.argmin(abs(band - redshifts))
                m = delta[i, zbin, :nkwf]
                after = np.dot(wf, m)
                if delta_err is not None:
                    up = np.dot(wf, m + delta_err[zbin, :nkwf])
                    lo = np.dot(wf, m - delta_err[zbin, :nkwf])
                    mean_err = np.mean([up - after, after - lo], axis=0)
                    err = np.sqrt(var + (0.2 * after) ** 2 + mean_err**2)
                else:
                    err = np.sqrt(var + (0.2 * after) ** 2)
                if gaussian:
                    lnl[i] += -0.5 * np.sum((after - vals) ** 2 / (err**2))
                    continue
                like = 0.5 + 0.5 * erf((vals - after) / (np.sqrt(2) * err))
                like[like <= 0.0] = 1e-50
                lnl[i] += np.nansum(np.log(like))
    return lnl


# --------------------------------------------------------------------------------------
# Independent cross-checks used by the tests (not reference restatements)
# --------------------------------------------------------------------------------------
def half_spectrum_power(box, boxlength, edges, weights_flags=(True, False, False), coeval=False):
    """Same spherical average from the Hermitian half-spectrum (rfftn) with multiplicity weights.

    SURVEY.md Appendix A step 9: this is the form the GPU path uses; the test-suite checks that it
    reproduces the full-grid ``get_power`` counts bit-exactly.
    """
    box = np.asarray(box, dtype=np.float64)
    n0, n1, n2 = box.shape
    L = np.array([float(boxlength)] * 3 if np.isscalar(boxlength) else [float(x) for x in boxlength])
    V = float(np.prod(L))
    F = np.fft.rfftn(box)
    P = (V / float(box.size) ** 2) * (F.real**2 + F.imag**2)
    k = [np.fft.fftfreq(n, d=l / n) * 2 * np.pi for n, l in zip(box.shape, L)]
    k2h = k[2][: n2 // 2 + 1].copy()
    if n2 % 2 == 0:
        k2h[-1] = abs(k2h[-1])
    kmag = np.sqrt((k[0][:, None, None] ** 2 + k[1][None, :, None] ** 2) + k2h[None, None, :] ** 2)
    mult = np.full(n2 // 2 + 1, 2.0)
    mult[0] = 1.0
    if n2 % 2 == 0:
        mult[-1] = 1.0
    w = np.ones(kmag.shape) * mult[None, None, :]
    ikperp, ikpar, ik0 = weights_flags
    if ikperp:
        w[0, 0, :] = 0
    if ikpar:
        w[:, :, 0] = 0
    if ik0:
        w[0, 0, 0] = 0
    indx = np.digitize(kmag.ravel(), edges)
    nb = len(edges) - 1
    sw = np.bincount(indx, weights=w.ravel(), minlength=nb + 2)[1:-1]
    res = np.bincount(indx, weights=(P * w).ravel(), minlength=nb + 2)[1:-1] / sw
    return res, sw
