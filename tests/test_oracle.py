"""Pins the CPU oracle (oracle/ps_oracle.py) with analytic known-answer tests.

The reference holds no golden vector for this path, so the oracle is pinned
by closed-form results (white noise, plane wave, Parseval), by internal consistency between the
full-grid restatement and the half-spectrum form the GPU uses, and by frozen vectors in
tests/golden/ (made by tests/golden/make_golden.py).
"""
import os

import numpy as np
import pytest

from oracle import ps_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_fftfreq_shift_zero_at_n_half():
    for n in (8, 9, 35, 128):
        f = O.fftfreq(n, d=1.5)
        assert f[n // 2] == 0.0
        assert np.all(np.diff(f) > 0)
        assert np.isclose(f[n // 2 + 1], 2 * np.pi / (n * 1.5))


def test_white_noise_kat():
    # <P_bin> = sigma^2 V / N for white Gaussian noise
    rng = np.random.default_rng(1)
    shape, L, sigma = (64, 64, 128), (96.0, 96.0, 192.0), 10.0
    box = (rng.standard_normal(shape) * sigma).astype(np.float32)
    p, k, sw = O.compute_power_lightcone(box, L, 12, log_bins=False, return_counts=True)
    expect = sigma**2 * np.prod(L) / np.prod(shape)
    good = sw > 2000
    assert good.sum() >= 6
    assert np.all(np.abs(p[good] / expect - 1) < 5 / np.sqrt(sw[good]) + 0.01)


def test_plane_wave_kat():
    # A cos(2 pi (a x/N0 + b y/N1 + c z/N2)): two modes of power V A^2/4 land in one bin
    n = (32, 32, 48)
    L = (64.0, 64.0, 96.0)
    A, (a, b, c) = 3.0, (3, 5, 7)
    x, y, z = np.meshgrid(*[np.arange(m) for m in n], indexing="ij")
    box = A * np.cos(2 * np.pi * (a * x / n[0] + b * y / n[1] + c * z / n[2]))
    p, k, sw = O.compute_power_lightcone(box, L, 10, log_bins=True, return_counts=True)
    kw = 2 * np.pi * np.sqrt((a / L[0]) ** 2 + (b / L[1]) ** 2 + (c / L[2]) ** 2)
    V = np.prod(L)
    hit = np.nanargmax(p)
    res, edges = O.get_power(box, L, bins=10, log_bins=True,
                             k_weights=O.k_weights_lightcone(n))
    assert edges[hit] <= kw < edges[hit + 1]
    assert np.isclose(p[hit], 2 * (V * A**2 / 4) / sw[hit], rtol=1e-10)
    others = np.delete(p, hit)
    assert np.all((np.abs(others[np.isfinite(others)]) < 1e-12 * p[hit]))


def test_parseval_linear_all_modes():
    # with every mode weighted once, sum_b P_b count_b = (V/N^2) N sum x^2
    rng = np.random.default_rng(2)
    shape, L = (16, 16, 20), (30.0, 30.0, 45.0)
    box = rng.standard_normal(shape)
    p, k, sw = O.compute_power_lightcone(box, L, 8, log_bins=False, ignore_kperp_zero=False,
                                         return_counts=True)
    assert sw.sum() == np.prod(shape)  # 1.001 top edge keeps every mode
    total = np.nansum(p * sw)
    assert np.isclose(total, np.prod(L) / np.prod(shape) * np.sum(box**2), rtol=1e-12)


def test_dc_power_is_kept_without_mean_subtraction():
    # gridded input is NOT mean-subtracted by get_power: P(0) = V mean^2 sits in linear bin 0
    shape, L = (8, 8, 8), 16.0
    box = np.full(shape, 2.5)
    p, k, sw = O.compute_power_coeval(box, L, 4, log_bins=False, ignore_kperp_zero=False,
                                      return_counts=True)
    assert np.isclose(p[0] * sw[0], L**3 * 2.5**2)


@pytest.mark.parametrize("shape,L,log", [
    ((32, 32, 64), (48.0, 48.0, 96.0), True),
    ((20, 20, 20), 30.0, False),
    ((15, 15, 17), (20.0, 20.0, 31.0), True),
    ((12, 12, 9), (18.0, 18.0, 13.5), False),
])
@pytest.mark.parametrize("rule", O.EDGE_RULES)
def test_half_spectrum_equals_full_grid(shape, L, log, rule):
    rng = np.random.default_rng(3)
    box = rng.standard_normal(shape) + 2.0
    for flags in ((True, False, False), (False, False, False), (True, True, False), (False, False, True)):
        p, k, sw = O.compute_power_lightcone(box, L if not np.isscalar(L) else (L, L, L), 9, log_bins=log,
                                             ignore_kperp_zero=flags[0], ignore_kpar_zero=flags[1],
                                             ignore_k_zero=flags[2], edge_rule=rule, return_counts=True)
        w = O.k_weights_lightcone(shape, *flags)
        _, edges = O.get_power(box, L, bins=9, log_bins=log, k_weights=w, edge_rule=rule)
        ph, swh = O.half_spectrum_power(box, L, edges, flags)
        assert np.array_equal(sw, swh)  # mode counts bit-exact
        m = sw > 0
        assert np.allclose(p[m], ph[m], rtol=1e-12, atol=0)
        assert np.all(np.isnan(p[~m])) and np.all(np.isnan(ph[~m]))


def test_empty_bins_are_nan_for_anisotropic_log_bins():
    # anisotropic chunks leave the lowest log bins empty -> NaN (kept)
    rng = np.random.default_rng(4)
    box = rng.standard_normal((32, 32, 64)).astype(np.float32)
    p, k, sw = O.compute_power_lightcone(box, (75.0, 75.0, 150.0), 20, return_counts=True)
    assert sw[0] == 0 and np.isnan(p[0])


def test_chunk_indices_follow_reference_rounding():
    assert O.chunk_indices(1024, 4) == [0, 256, 512, 768, 1024]
    assert O.chunk_indices(1000, 3) == [0, 333, 666, 1000]
    assert O.chunk_indices(10, 4) == [0, 2, 4, 6, 10]  # round(2.5) == 2 (banker's)
    assert O.chunk_indices(7, 2) == [0, 4, 7]
    with pytest.raises(IndexError):
        O.reduce_data_lightcone(np.zeros((4, 4, 6), np.float32), 6.0, 1.5, nchunks=4, n_psbins=3)


def test_default_bins_and_centres():
    rng = np.random.default_rng(5)
    box = rng.standard_normal((20, 20, 20))
    p, k = O.compute_power_coeval(box, 30.0, None)
    assert len(p) == int(20 / 2.2) and len(k) == len(p)
    _, edges = O.get_power(box, 30.0, bins=None, log_bins=True, k_weights=O.k_weights_coeval(box.shape))
    assert np.allclose(k, np.sqrt(edges[1:] * edges[:-1]))


def test_cylindrical_mirror_symmetry():
    rng = np.random.default_rng(6)
    box = rng.standard_normal((16, 16, 12))
    P2, edges, kpar = O.get_power(box, (24.0, 24.0, 18.0), bins=5, res_ndim=2)
    assert P2.shape == (5, 12) and len(kpar[0]) == 12
    # k_par shifted: index 6 is 0; +/- mirror columns are equal for a real field
    for j in range(1, 6):
        assert np.allclose(P2[:, 6 + j], P2[:, 6 - j], rtol=1e-12, equal_nan=True)


def test_likelihood_zero_on_own_data_and_chi2_value():
    rng = np.random.default_rng(7)
    lc = (rng.standard_normal((24, 24, 96)) * 10 + 20).astype(np.float32)
    data = O.reduce_data_lightcone(lc, 36.0, 1.5, nchunks=3, n_psbins=8)
    assert O.compute_likelihood(data, data, min_k=0.2, max_k=2.0) == 0
    model = [{"k": d["k"], "delta": d["delta"] * 1.3} for d in data]
    lnl = O.compute_likelihood(model, data, min_k=0.2, max_k=2.0, model_uncertainty=0.15)
    nsel = sum(int(np.sum((d["k"] >= 0.2) & (d["k"] <= 2.0) & np.isfinite(d["delta"]))) for d in data)
    assert np.isclose(lnl, -0.5 * nsel * (0.3 / (0.15 * 1.3)) ** 2, rtol=1e-10)


def test_float32_fft_mode_close_to_float64():
    rng = np.random.default_rng(8)
    box = (rng.standard_normal((32, 32, 64)) * 10 + 20).astype(np.float32)
    p64, _ = O.compute_power_lightcone(box, (48.0, 48.0, 96.0), 10)
    p32, _ = O.compute_power_lightcone(box, (48.0, 48.0, 96.0), 10, fft_dtype=np.float32)
    m = np.isfinite(p64)
    assert np.max(np.abs(p32[m] / p64[m] - 1)) < 2e-6


def test_golden_vectors():
    gold = np.load(os.path.join(GOLD, "oracle_golden.npz"), allow_pickle=False)
    from tests.golden.make_golden import CASES, make_box

    for name, case in CASES.items():
        box = make_box(case)
        kw = dict(log_bins=case["log"], ignore_kperp_zero=case["flags"][0], ignore_kpar_zero=case["flags"][1],
                  ignore_k_zero=case["flags"][2], edge_rule=case["rule"], return_counts=True)
        p, k, sw = O.compute_power_lightcone(box, case["L"], case["nb"], **kw)
        assert np.array_equal(sw.astype(np.int64), gold[f"{name}_counts"])
        assert np.array_equal(k, gold[f"{name}_k"])
        assert np.allclose(p, gold[f"{name}_power"], rtol=1e-12, atol=0, equal_nan=True)
