"""Seeded random sweep of the real library against the oracle (run on a GPU box):
general-path shapes incl. primes and odd sizes, spherical and cylindrical, both edge rules, all ignore flags.
usage: python scripts/gpu_random_sweep.py [ncases] [seed] [smax] [nzmax]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import py21cmmc_b200  # noqa: E402,F401
from py21cmmc_b200 import likelihood as lk  # noqa: E402
from tests import parity_cases as PC  # noqa: E402


# No tolerance floor: every finite bin is held to the 1e-5 contract.  What the sweep avoids instead are output cells
# that consist of ONE Fourier mode: an fp32 FFT (this library's, and NumPy's float32 FFT, i.e. the reference's own
# numerics) resolves a mode to ~1e-7 of the RMS amplitude, so a single mode that happens to carry 1e-4 of the typical
# power is already past 1e-5 (seen on an 8 x 8 x 142 cylindrical case: that cell 3.9e-5 here, 7e-6 in NumPy float32).
# Hence sky planes start at 8 x 8 cells, and cylindrical (k_perp, k_par) cases need a 16 x 16 plane and at most
# HII_DIM/3 k_perp bins, which puts at least four modes into every cell.
FLOOR = 0.0


def sweep(n, seed, lk_mod=lk, smax=72, nzmax=150):
    """Returns (number of failing cases, worst relative power error of the passing ones)."""
    rng = np.random.default_rng(seed)
    worst, fails = 0.0, 0
    for i in range(n):
        s = int(rng.integers(8, smax))
        nz = int(rng.integers(4, nzmax))
        coeval = rng.random() < 0.25
        shape = (s, s, s) if coeval else (s, s, nz)
        cell = float(rng.uniform(0.5, 3.0))
        L = s * cell if coeval else (s * cell, s * cell, nz * cell)
        nb = None if rng.random() < 0.2 else int(rng.integers(2, 60))
        log = bool(rng.random() < 0.6)
        flags = tuple(bool(rng.random() < p) for p in (0.7, 0.3, 0.3))
        rule = "legacy_1.001" if rng.random() < 0.6 else "inscribed"
        try:
            if rng.random() < 0.2 and not coeval and s >= 16:
                err = PC.check_cylindrical(lk_mod, shape, L, int(rng.integers(2, max(3, min(12, s // 3)))), seed=i,
                                           floor_frac=FLOOR)
            else:
                err = PC.check_compute_power(lk_mod, shape, L, nb, log, flags, rule, seed=i, coeval=coeval,
                                             floor_frac=FLOOR)
            worst = max(worst, err)
        except Exception as e:  # noqa: BLE001
            fails += 1
            print("FAIL", i, shape, L, nb, log, flags, rule, coeval, repr(e)[:200])
    return fails, worst


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 150
    fails, worst = sweep(n, int(sys.argv[2]) if len(sys.argv) > 2 else 2024,
                         smax=int(sys.argv[3]) if len(sys.argv) > 3 else 72,
                         nzmax=int(sys.argv[4]) if len(sys.argv) > 4 else 150)
    print(f"random sweep: {n} cases, {fails} failures, worst power rel err {worst:.3e}")
    sys.exit(1 if fails else 0)
