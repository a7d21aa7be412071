"""Hypothesis sweep: random small shapes / lengths / options through the emulated kernels vs the oracle.

Covers what the fixed cases cannot enumerate: every mix of even/odd/prime axis lengths, anisotropic
boxes, both bin spacings, both top-edge rules, all ignore_* combinations, default bin counts.
"""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

from oracle import ps_oracle as O
from tests.helpers import POWER_RTOL, rel_err, synth_box


@pytest.fixture(scope="module")
def lk():
    from tests.emu.build_emu import build

    lib = build()
    import py21cmmc_b200  # noqa: F401
    from py21cmmc_b200 import _capi, likelihood, plan

    _capi._testing_set_library(lib)
    yield likelihood
    plan.clear_plans()
    _capi._testing_set_library(None)


@settings(max_examples=40, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(n=st.integers(4, 14), nz=st.integers(4, 40), cell=st.sampled_from([1.0, 1.5, 2.0]),
       nb=st.one_of(st.none(), st.integers(2, 12)), log=st.booleans(),
       flags=st.tuples(st.booleans(), st.booleans(), st.booleans()), rule=st.sampled_from(O.EDGE_RULES),
       seed=st.integers(0, 5))
def test_power_matches_oracle(lk, n, nz, cell, nb, log, flags, rule, seed):
    shape = (n, n, nz)
    L = (n * cell, n * cell, nz * cell)
    box = synth_box(shape, seed)
    if nb is None and int((n * n * nz) ** (1 / 3) / 2.2) < 1:
        return  # powerbox's default would be zero bins
    lk.set_defaults(edge_rule=rule)
    try:
        p, k = lk.Likelihood1DPowerLightcone.compute_power(box, L, nb, log, *flags)
    finally:
        lk.set_defaults(edge_rule="legacy_1.001")
    po, ko, sw = O.compute_power_lightcone(box, L, nb, log, *flags, edge_rule=rule, return_counts=True)
    assert np.array_equal(k, ko)
    from py21cmmc_b200.plan import get_plan

    plan = get_plan(shape, [(0, nz)], [L], len(po), log, *flags, rule, "spherical", lk._device(None), "auto")
    assert np.array_equal(plan.counts, sw.astype(np.int64))
    assert rel_err(p, po) < POWER_RTOL


@settings(max_examples=15, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(n=st.integers(4, 10), nslices=st.integers(8, 60), nchunks=st.integers(1, 5), nb=st.integers(2, 8),
       log=st.booleans())
def test_lightcone_chunks_match_oracle(lk, n, nslices, nchunks, nb, log):
    from py21cmmc_b200 import binning
    from py21cmmc_b200.plan import get_plan

    ci = binning.chunk_indices(nslices, nchunks)
    if len(ci) < nchunks + 1:
        with pytest.raises(IndexError):
            O.reduce_data_lightcone(np.zeros((n, n, nslices), np.float32), n * 1.5, 1.5, nchunks, nb, log)
        return
    assert ci == O.chunk_indices(nslices, nchunks)
    shape = (n, n, nslices)
    box = synth_box(shape, 1)
    bounds = [(ci[i], ci[i + 1]) for i in range(nchunks)]
    lengths = [(n * 1.5, n * 1.5, (b - a) * 1.5) for a, b in bounds]
    if any(b - a < 2 for a, b in bounds):
        return
    plan = get_plan(shape, bounds, lengths, nb, log, True, False, False, "legacy_1.001", "spherical",
                    lk._device(None), "auto")
    rec = plan.power_host(box)[0]
    mo = O.reduce_data_lightcone(box, n * 1.5, 1.5, nchunks, nb, log)
    for c in range(nchunks):
        k = plan.kcentres[c]
        assert np.array_equal(k, mo[c]["k"])
        assert rel_err(plan._chunk_view(rec, c) * k**3 / (2 * np.pi**2), mo[c]["delta"]) < POWER_RTOL
