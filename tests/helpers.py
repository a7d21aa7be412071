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
