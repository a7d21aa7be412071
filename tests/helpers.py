"""Shared test helpers: deterministic synthetic boxes and oracle-side evaluation."""
import numpy as np

from tests.golden.make_golden import make_box


def synth_box(shape, seed=0):
    """float32 box, mean ~20, sigma ~10 (the benchmark's statistics), from an integer hash."""
    return make_box({"shape": tuple(shape)}, seed=seed)


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    m = np.isfinite(b)
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN (empty-bin) pattern differs"
    if not m.any():
        return 0.0
    return float(np.max(np.abs(a[m] - b[m]) / np.abs(b[m])))


# tolerances stated by BASELINE.json north_star
POWER_RTOL = 1e-5
LNL_RTOL = 1e-6
