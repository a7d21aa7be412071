"""CPU parity: the *same* kernel sources as libps21.so (general path), built by g++ against the host
CUDA emulation in tests/emu, driven through the C ABI and the plugin classes, versus the oracle.
This checks index maths, FFT stages, binning, plan bookkeeping and the host layer without a GPU; the
`-m gpu` tests repeat it on the real library."""
import numpy as np
import pytest

from tests import parity_cases as PC


@pytest.fixture(scope="module")
def mods():
    from tests.emu.build_emu import build

    lib = build()
    import py21cmmc_b200
    from py21cmmc_b200 import _capi, chain, core, likelihood, plan

    _capi._testing_set_library(lib)
    yield likelihood, core, chain
    plan.clear_plans()
    _capi._testing_set_library(None)


@pytest.mark.parametrize("shape,L,nb,log,flags,rule", [
    ((16, 16, 32), (24.0, 24.0, 48.0), 8, True, (True, False, False), "legacy_1.001"),
    ((16, 16, 32), (24.0, 24.0, 48.0), 8, False, (False, False, False), "legacy_1.001"),  # DC binned
    ((12, 12, 37), (18.0, 18.0, 55.5), 6, True, (True, False, False), "legacy_1.001"),   # prime length
    ((15, 15, 17), (22.5, 22.5, 25.5), 7, True, (True, True, False), "legacy_1.001"),    # odd sizes
    ((24, 24, 18), (36.0, 36.0, 27.0), 10, False, (True, False, True), "inscribed"),
    ((10, 10, 14), (15.0, 15.0, 21.0), None, True, (True, False, False), "legacy_1.001"),  # default bins
    ((512, 512, 4), (768.0, 768.0, 6.0), 10, True, (True, False, False), "legacy_1.001"),  # radix-32 stages (y, x)
    ((100, 100, 6), (150.0, 150.0, 9.0), 12, False, (True, False, False), "legacy_1.001"),  # radix-25 stages (y, x)
    ((75, 75, 10), (110.0, 110.0, 15.0), 9, True, (True, False, False), "inscribed"),      # 75 = 25 * 3
])
def test_compute_power_lightcone(mods, shape, L, nb, log, flags, rule):
    PC.check_compute_power(mods[0], shape, L, nb, log, flags, rule)


def test_compute_power_coeval(mods):
    PC.check_compute_power(mods[0], (20, 20, 20), 30.0, 9, True, coeval=True)
    with pytest.raises(ValueError):
        mods[0].Likelihood1DPowerCoeval.compute_power(np.zeros((8, 8, 6), np.float32), 10.0, 4)


def test_lightcone_likelihood_chain(mods):
    PC.check_lightcone_likelihood(*mods, shape=(16, 16, 48), box_len=24.0, nchunks=3, nb=8, min_k=0.3, max_k=3.0)


def test_lightcone_unequal_chunks_linear(mods):
    # 50 slices / 3 chunks -> 17, 17, 16: two shapes in one plan
    PC.check_lightcone_likelihood(*mods, shape=(12, 12, 50), box_len=18.0, nchunks=3, nb=6, min_k=0.2, max_k=4.0,
                                  logk=False)


def test_coeval_likelihood_chain(mods):
    PC.check_coeval_likelihood(*mods, n=12, box_len=18.0, nb=6)


def test_cylindrical(mods):
    PC.check_cylindrical(mods[0], (12, 12, 16), (18.0, 18.0, 24.0), 5)
    PC.check_cylindrical(mods[0], (10, 10, 9), (15.0, 15.0, 13.5), 4)


@pytest.mark.parametrize("nb", [20, None])   # None: the reference default, 116 bins -> 4-lane group bin mode
def test_fused_two_pass_kernels_256(mods, nb):
    """The fused sm_100a kernels (bulk copies modelled as memcpy) on one 256^3 chunk vs the oracle."""
    mods[0].set_defaults(kernel_path="fused")
    try:
        PC.check_compute_power(mods[0], (256, 256, 256), (384.0, 384.0, 384.0), nb, True, seed=5)
    finally:
        mods[0].set_defaults(kernel_path="auto")


@pytest.mark.parametrize("nb", [20, None])   # None: the reference default, 73 bins -> 4-lane group bin mode
def test_fused_two_pass_kernels_128(mods, nb):
    """128 x 128 x 256 chunk (BASELINE config 1 chunk shape) through the fused kernels."""
    mods[0].set_defaults(kernel_path="fused")
    try:
        PC.check_compute_power(mods[0], (128, 128, 256), (192.0, 192.0, 384.0), nb, True, seed=3)
    finally:
        mods[0].set_defaults(kernel_path="auto")


def test_fused_cylindrical_256(mods):
    """Cylindrical (k_perp, k_par) binning inside the fused pass 2 (BASELINE config 5 chunk shape)."""
    mods[0].set_defaults(kernel_path="fused")
    try:
        PC.check_cylindrical(mods[0], (256, 256, 256), (384.0, 384.0, 384.0), 12, seed=2)
    finally:
        mods[0].set_defaults(kernel_path="auto")


def test_upper_limit_likelihood(mods, tmp_path):
    PC.check_upper_limit_likelihood(*mods, tmp_path)


def test_too_few_slices_raises_like_reference(mods):
    lk, core, chain = mods
    with pytest.raises(IndexError):
        PC.make_lightcone_chain(lk, core, chain, (4, 4, 6), 6.0, nchunks=4, nb=3)


def test_lnl_before_data_is_a_state_error(mods):
    from py21cmmc_b200 import Ps21Error
    from py21cmmc_b200.plan import PowerPlan
    from py21cmmc_b200 import _capi

    pl = PowerPlan((8, 8, 8), [(0, 8)], [(12.0, 12.0, 12.0)], 4, library=_capi.LIB_PATH)
    with pytest.raises(Ps21Error):
        pl.lnl_host(np.zeros((1, 8, 8, 8), np.float32))
    with pytest.raises(ValueError):
        pl.power_host(np.zeros((1, 8, 8, 4), np.float32))
    pl.close()


@pytest.mark.parametrize("shape,L,nb,log,flags", [
    ((16, 16, 32), (24.0, 24.0, 48.0), 8, True, (True, False, False)),
    ((32, 32, 20), (48.0, 48.0, 30.0), 6, False, (False, False, False)),   # DC binned
    ((40, 40, 40), (60.0, 60.0, 60.0), 9, True, (True, True, False)),      # 40 = 8 * 5
    ((24, 24, 28), (36.0, 36.0, 42.0), 7, True, (True, False, True)),      # nz / 2 = 14 = 2 * 7
    ((48, 48, 52), (72.0, 72.0, 78.0), None, True, (True, False, False)),  # 2 * 13
    ((80, 80, 44), (120.0, 120.0, 66.0), 50, False, (True, False, False)), # 2 * 11, 50 bins: 4-lane-group bin columns
    ((64, 64, 16), (96.0, 96.0, 24.0), 10, True, (True, False, False)),    # R2 = 4, x = 8 * 8
    ((200, 200, 8), (300.0, 300.0, 12.0), 20, True, (True, False, False)), # config-2 sky plane: y 50 x 4, x = 8 * 25
    ((16, 16, 34), (24.0, 24.0, 51.0), 8, True, (True, False, False)),     # nz / 2 = 17: Bluestein (chirp convolution)
    ((32, 32, 38), (48.0, 48.0, 57.0), 6, False, (False, False, False)),   # 19, DC binned
    ((64, 64, 334), (96.0, 96.0, 501.0), 12, True, (True, False, False)),  # 167: the 334-slice chunk of a 1002-slice lightcone
    ((16, 16, 15), (24.0, 24.0, 22.5), 8, True, (True, False, False)),     # odd chunk length: transform of row pairs
    ((32, 32, 21), (48.0, 48.0, 31.5), 6, False, (False, False, False)),   # odd, 3 * 7, DC binned
    ((24, 24, 37), (36.0, 36.0, 55.5), 7, True, (True, False, True)),      # odd and prime: row pairs + Bluestein
    ((64, 64, 333), (96.0, 96.0, 499.5), 12, True, (True, False, False)),  # 333 = 9 * 37: the other chunks of that lightcone
])
def test_generic_two_pass_kernels(mods, shape, L, nb, log, flags):
    """ps21_g2.cu (in-place digit-reversed transforms, position-defined tiles) on the emulated device vs the oracle."""
    from py21cmmc_b200.plan import get_plan

    mods[0].set_defaults(kernel_path="twopass")
    try:
        PC.check_compute_power(mods[0], shape, L, nb, log, flags)
    finally:
        mods[0].set_defaults(kernel_path="auto")


def test_two_pass_rejects_what_it_cannot_factor(mods):
    from py21cmmc_b200 import Ps21Error
    from py21cmmc_b200.plan import PowerPlan
    from py21cmmc_b200 import _capi

    for shape in ((15, 15, 16), (50, 50, 16)):   # odd sky plane; 50 rows do not split into 4 or 8 units
        with pytest.raises(Ps21Error):
            PowerPlan(shape, [(0, shape[2])], [(24.0, 24.0, 30.0)], 4, path="twopass", library=_capi.LIB_PATH)
    pl = PowerPlan((15, 15, 16), [(0, 16)], [(22.5, 22.5, 24.0)], 4, path="auto", library=_capi.LIB_PATH)
    assert pl.path == "general"
    pl.close()
