"""Parity checks shared by the CPU (emulated kernels) and GPU (libps21.so) test modules.

Every function compares the product path, called through the C ABI / the plugin classes, with
oracle/ps_oracle.py on the same inputs.  Tolerances are BASELINE.json's: mode counts and bin
assignment bit-exact, power 1e-5 relative, lnL 1e-6 relative.
"""
import numpy as np

from oracle import ps_oracle as O
from tests.helpers import LNL_RTOL, POWER_RTOL, rel_err, synth_box


def check_compute_power(lk_mod, shape, L, nb, log, flags=(True, False, False), rule="legacy_1.001", seed=0,
                        coeval=False):
    from py21cmmc_b200.plan import get_plan

    box = synth_box(shape, seed)
    lk_mod.set_defaults(edge_rule=rule)
    try:
        if coeval:
            p, k = lk_mod.Likelihood1DPowerCoeval.compute_power(box, L, nb, log, *flags)
            po, ko, sw = O.compute_power_coeval(box, L, nb, log, *flags, edge_rule=rule, return_counts=True)
            Ls = (L, L, L)
        else:
            p, k = lk_mod.Likelihood1DPowerLightcone.compute_power(box, L, nb, log, *flags)
            po, ko, sw = O.compute_power_lightcone(box, L, nb, log, *flags, edge_rule=rule, return_counts=True)
            Ls = L
        plan = get_plan(shape, [(0, shape[2])], [Ls], len(po), log, *flags, rule, "spherical",
                        lk_mod._device(None), lk_mod._DEFAULTS["kernel_path"])
    finally:
        lk_mod.set_defaults(edge_rule="legacy_1.001")
    assert np.array_equal(k, ko), "bin centres must be bit-identical (same NumPy expressions)"
    assert np.array_equal(plan.counts, sw.astype(np.int64)), "mode counts must be bit-exact"
    err = rel_err(p, po)
    assert err < POWER_RTOL, f"power rel err {err}"
    return err


def make_lightcone_chain(lk_mod, core_mod, chain_mod, shape, box_len, nchunks, nb, seed=0, simulate=True,
                         **lk_kwargs):
    lc = synth_box(shape, seed)
    up = core_mod.UserParams(BOX_LEN=box_len, HII_DIM=shape[0])
    state = {"box": lc}
    c = core_mod.CoreLightConeModule(lambda params: state["box"], up)
    lk = lk_mod.Likelihood1DPowerLightcone(nchunks=nchunks, n_psbins=nb, simulate=simulate, **lk_kwargs)
    chain = chain_mod.build_computation_chain(c, lk, setup=True)
    return chain, lk, state


def check_lightcone_likelihood(lk_mod, core_mod, chain_mod, shape, box_len, nchunks, nb, min_k, max_k,
                               logk=True):
    chain, lk, state = make_lightcone_chain(lk_mod, core_mod, chain_mod, shape, box_len, nchunks, nb,
                                            min_k=min_k, max_k=max_k, logk=logk)
    cell = box_len / shape[0]
    data_o = O.reduce_data_lightcone(state["box"], box_len, cell, nchunks, nb, logk)
    # the mock data made at setup() through the device path equals the oracle's
    for d, do in zip(lk.data, data_o):
        assert np.array_equal(d["k"], do["k"])
        assert rel_err(d["delta"], do["delta"]) < POWER_RTOL
    # a different "walker"
    state["box"] = synth_box(shape, seed=7)
    ctx = chain.build_model_data()
    model = lk.reduce_data(ctx)
    model_o = O.reduce_data_lightcone(state["box"], box_len, cell, nchunks, nb, logk)
    for m, mo in zip(model, model_o):
        assert rel_err(m["delta"], mo["delta"]) < POWER_RTOL
    lnl = lk.computeLikelihood(model)
    lnl_o = O.compute_likelihood(model_o, data_o, None, min_k, max_k, 0.15, True)
    assert np.isfinite(lnl_o) and lnl_o < 0
    assert abs(lnl - lnl_o) <= LNL_RTOL * abs(lnl_o), (lnl, lnl_o)
    # store(): blob keys of the reference
    blobs = {}
    lk.store(model, blobs)
    assert "k_0" in blobs and "delta_0" in blobs and f"delta_{nchunks - 1}" in blobs
    # batched walkers == per-walker path, and the fused device chi^2 agrees with the oracle
    boxes = [synth_box(shape, seed=s) for s in (7, 8, 9)]
    ctxs = []
    for b in boxes:
        state["box"] = b
        ctxs.append(chain.build_model_data())
    lnls, models = lk.likelihood_batch(ctxs)
    assert abs(lnls[0] - lnl_o) <= LNL_RTOL * abs(lnl_o)
    for b, l, ms in zip(boxes, lnls, models):
        mo = O.reduce_data_lightcone(b, box_len, cell, nchunks, nb, logk)
        lo = O.compute_likelihood(mo, data_o, None, min_k, max_k, 0.15, True)
        assert abs(l - lo) <= LNL_RTOL * abs(lo)
        for m, mm in zip(ms, mo):
            assert rel_err(m["delta"], mm["delta"]) < POWER_RTOL
    return lnl, lnl_o


def check_coeval_likelihood(lk_mod, core_mod, chain_mod, n, box_len, nb, redshifts=(7.0, 8.0, 9.0)):
    up = core_mod.UserParams(BOX_LEN=box_len, HII_DIM=n)
    state = {"seed": 0}

    def sim(params):
        return [synth_box((n, n, n), seed=state["seed"] * 10 + i) for i in range(len(redshifts))]

    c = core_mod.CoreCoevalModule(list(redshifts), sim, up)
    lk = lk_mod.Likelihood1DPowerCoeval(n_psbins=nb, simulate=True, min_k=0.3, max_k=3.0)
    chain = chain_mod.build_computation_chain(c, lk, setup=True)
    data_o = O.reduce_data_coeval(sim(None), box_len, nb)
    state["seed"] = 3
    ctx = chain.build_model_data()
    model = lk.reduce_data(ctx)
    model_o = O.reduce_data_coeval(sim(None), box_len, nb)
    assert len(model) == len(redshifts)
    for m, mo in zip(model, model_o):
        assert np.array_equal(m["k"], mo["k"])
        assert rel_err(m["delta"], mo["delta"]) < POWER_RTOL
    lnl = lk.computeLikelihood(model)
    lnl_o = O.compute_likelihood(model_o, data_o, None, 0.3, 3.0, 0.15, True)
    assert abs(lnl - lnl_o) <= LNL_RTOL * abs(lnl_o)
    blobs = {}
    lk.store(model, blobs)
    assert "k_z7.0" in blobs and "delta_z9.0" in blobs
    lnls, _ = lk.likelihood_batch([ctx, ctx])
    assert np.allclose(lnls, lnl_o, rtol=LNL_RTOL, atol=0)
    # the chain end to end
    val, blob = chain([])
    assert abs(val - lnl_o) <= LNL_RTOL * abs(lnl_o) and "delta_z8.0" in blob
    return lnl, lnl_o


def check_cylindrical(lk_mod, shape, L, nb, seed=0):
    box = synth_box(shape, seed)
    P2, kperp, kpar = lk_mod.Likelihood1DPowerLightcone.compute_power_2d(box, L, nb)
    w2 = np.ones(shape[:2])
    w2[shape[0] // 2, shape[0] // 2] = 0
    Po, edges, kpo = O.get_power(box, L, bins=nb, res_ndim=2, k_weights=w2)
    assert P2.shape == Po.shape
    assert np.array_equal(kperp, O.edges_to_centres(edges, False))
    assert np.allclose(kpar, kpo[0], rtol=1e-15)
    assert rel_err(P2, Po) < POWER_RTOL
