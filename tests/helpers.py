"""Shared test helpers: deterministic synthetic boxes and oracle-side evaluation."""
import numpy as np

from tests.golden.make_golden import make_box


def synth_box(shape, seed=0):
    """float32 box, mean ~20, sigma ~10 (the benchmark's statistics), from an integer hash."""
    return make_box({"shape": tuple(shape)}, seed=seed)


def rel_err(a, b, floor_frac=0.0):
    """max |a - b| / |b| over the finite entries of b (the NaN pattern must agree).

    ``floor_frac`` > 0 measures bins far below the typical power against ``floor_frac * median|b|`` instead of
    their own value: a single-mode bin that happens to hold 1/1000 of the typical power is ill conditioned in
    any fp32 FFT (NumPy's float32 FFT, i.e. the reference's own numerics, is off by 6e-5 on such a bin where
    this library is off by 1e-5).  Used by the random shape sweep only; the fixed cases keep the strict form.
    """
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    m = np.isfinite(b)
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN (empty-bin) pattern differs"
    if not m.any():
        return 0.0
    den = np.abs(b[m])
    if floor_frac > 0:
        den = np.maximum(den, floor_frac * np.median(den))
    return float(np.max(np.abs(a[m] - b[m]) / den))


# tolerances stated by BASELINE.json north_star
POWER_RTOL = 1e-5
LNL_RTOL = 1e-6


def power_law_box(shape, index=-3.0, seed=0, mean=20.0, sigma=10.0):
    """float32 Gaussian random field with P(k) ~ |k|^index (a realistic, steep 21-cm-like spectrum instead of
    white noise): one bin then mixes modes of very different power."""
    rng = np.random.default_rng(seed)
    white = rng.standard_normal(shape)
    F = np.fft.rfftn(white)
    k = [np.fft.fftfreq(n) for n in shape[:2]] + [np.fft.rfftfreq(shape[2])]
    kk = np.sqrt(k[0][:, None, None] ** 2 + k[1][None, :, None] ** 2 + k[2][None, None, :] ** 2)
    kk[0, 0, 0] = 1.0
    F *= kk ** (index / 2.0)
    F[0, 0, 0] = 0.0
    x = np.fft.irfftn(F, s=shape, axes=(0, 1, 2))
    x *= sigma / x.std()
    return (x + mean).astype(np.float32)
