"""Parity checks shared by the CPU (emulated kernels) and GPU (libps21.so) test modules.

Every function compares the product path, called through the C ABI / the plugin classes, with
oracle/ps_oracle.py on the same inputs.  Tolerances are BASELINE.json's: mode counts and bin
assignment bit-exact, power 1e-5 relative, lnL 1e-6 relative.
"""
import numpy as np

from oracle import ps_oracle as O
from tests.helpers import LNL_RTOL, POWER_RTOL, rel_err, synth_box


def check_compute_power(lk_mod, shape, L, nb, log, flags=(True, False, False), rule="legacy_1.001", seed=0,
                        coeval=False, floor_frac=0.0):
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
    err = rel_err(p, po, floor_frac)
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


def check_cylindrical(lk_mod, shape, L, nb, seed=0, floor_frac=0.0):
    box = synth_box(shape, seed)
    P2, kperp, kpar = lk_mod.Likelihood1DPowerLightcone.compute_power_2d(box, L, nb)
    w2 = np.ones(shape[:2])
    w2[shape[0] // 2, shape[0] // 2] = 0
    Po, edges, kpo = O.get_power(box, L, bins=nb, res_ndim=2, k_weights=w2)
    assert P2.shape == Po.shape
    assert np.array_equal(kperp, O.edges_to_centres(edges, False))
    assert np.allclose(kpar, kpo[0], rtol=1e-15)
    err = rel_err(P2, Po, floor_frac)
    assert err < POWER_RTOL, f"cylindrical power rel err {err}"
    return err


def synthetic_hera_file(path, kmodel, seed=0):
    """A small file in the layout of src/py21cmmc/data/HERA_H1C_IDR3.npz (2 bands, 3 fields, NaN-padded k)."""
    rng = np.random.default_rng(seed)
    out = {"z_bands": np.array([10.4, 7.9])}
    for z, nkwf in ((10, 12), (8, 9)):
        kwf = np.linspace(kmodel.min() * 1.2, kmodel.max() * 0.6, nkwf)
        nk = 7
        band = np.full((3, nk, 3), np.nan)
        wf = np.zeros((3, nk, nkwf))
        for f in range(3):
            used = nk - f  # ragged: trailing k entries are NaN
            band[f, :used, 0] = np.linspace(kwf[1], kwf[-2], used)
            band[f, :used, 1] = rng.uniform(50.0, 4000.0, used)       # upper limits, mK^2
            band[f, :used, 2] = rng.uniform(10.0, 900.0, used) ** 2    # variances
            w = rng.uniform(0.0, 1.0, (used, nkwf))
            wf[f, :used] = w / w.sum(axis=1, keepdims=True)
        out[f"band{z}"], out[f"wfband{z}"], out[f"kwfband{z}"] = band, wf, kwf
    np.savez(path, **out)
    return out


def check_upper_limit_likelihood(lk_mod, core_mod, chain_mod, tmp_path, shape=(16, 16, 48), box_len=24.0):
    nchunks, nb = 3, 8
    up = core_mod.UserParams(BOX_LEN=box_len, HII_DIM=shape[0])
    state = {"box": synth_box(shape, 1)}
    c = core_mod.CoreLightConeModule(lambda params: state["box"], up)
    # the k grid of the chunks decides a sensible window-function k range
    probe = lk_mod.Likelihood1DPowerLightcone(nchunks=nchunks, n_psbins=nb, use_data=False)
    ch0 = chain_mod.build_computation_chain(c, probe)
    kmodel = probe.reduce_data(ch0.build_model_data())[0]["k"]
    hera = synthetic_hera_file(str(tmp_path / "hera.npz"), kmodel)
    lk = lk_mod.Likelihood1DPowerLightconeUpper(datafile=str(tmp_path / "hera.npz"), nchunks=nchunks, n_psbins=nb,
                                                band_chunks=[2, 0])
    chain = chain_mod.build_computation_chain(core_mod.CoreLightConeModule(lambda params: state["box"], up), lk)
    model = lk.reduce_data(chain.build_model_data())
    assert model[0]["delta"].shape == (1, 2, 12)
    lnl = lk.computeLikelihood(model)
    # oracle: chunk spectra from the oracle, same interpolation, reference arithmetic
    from scipy.interpolate import InterpolatedUnivariateSpline

    mo = O.reduce_data_lightcone(state["box"], box_len, box_len / shape[0], nchunks, nb)
    delta_o = np.zeros((1, 2, 12))
    for b, cidx in enumerate([2, 0]):
        ok = np.isfinite(mo[cidx]["delta"])
        kk = hera["kwfband" + ("10" if b == 0 else "8")]
        delta_o[0, b, :len(kk)] = InterpolatedUnivariateSpline(mo[cidx]["k"][ok], mo[cidx]["delta"][ok], k=1)(kk)
    lnl_o = O.upper_limit_likelihood(hera, hera["z_bands"], delta_o)[0]
    assert np.isfinite(lnl_o) and lnl_o < 0
    assert abs(lnl - lnl_o) <= 1e-6 * abs(lnl_o), (lnl, lnl_o)
    # the device kernel alone, many models + an error term, against the reference arithmetic
    rng = np.random.default_rng(5)
    many = {"k": lk.k, "delta": delta_o * rng.uniform(0.2, 30.0, (9, 1, 1)), "delta_err": np.abs(delta_o[0]) * 0.1}
    got = lk.computeLikelihood([many])
    want = O.upper_limit_likelihood(hera, hera["z_bands"], many["delta"], many["delta_err"])
    assert np.allclose(got, want, rtol=1e-10, atol=1e-12)
    val, blobs = chain([])
    assert abs(val - lnl_o) <= 1e-6 * abs(lnl_o) and "delta_0" in blobs
    # measured band powers (Gaussian branch, likelihood.py:517-583) through the same window functions
    lkg = lk_mod.Likelihood1DPowerLightconeWindowed(datafile=str(tmp_path / "hera.npz"), nchunks=nchunks,
                                                    n_psbins=nb, band_chunks=[2, 0])
    chain_mod.build_computation_chain(core_mod.CoreLightConeModule(lambda params: state["box"], up), lkg)
    gotg = lkg.computeLikelihood([many])
    wantg = O.upper_limit_likelihood(hera, hera["z_bands"], many["delta"], many["delta_err"], gaussian=True)
    assert np.all(wantg < 0) and np.allclose(gotg, wantg, rtol=1e-10, atol=1e-12)
    return lnl, lnl_o


def check_noise_and_error_on_model(mods, tmp_path, error_on_model, shape, box_len, nchunks, nb, min_k=0.3, max_k=4.0):
    """Noise spline, error_on_model on / off and data on another k grid than the model (likelihood.py:584-604):
    per-model chi^2 entry and the fused finalize kernel against the oracle."""
    lk_mod, core_mod, chain_mod = mods
    state = {"box": synth_box(shape, 0)}
    cell = box_len / shape[0]
    data_o = O.reduce_data_lightcone(state["box"], box_len, cell, nchunks, nb)
    dfiles, nfiles, noise = [], [], []
    rng = np.random.default_rng(1)
    for i, d in enumerate(data_o):
        ok = np.isfinite(d["delta"])
        f = str(tmp_path / f"d{i}.npz")
        np.savez(f, k=d["k"][ok], delta=d["delta"][ok] * 2.5)  # data on a *different* k grid than the model; far enough
        # from the model that the chi^2 does not amplify the fp32-FFT error of Delta^2 (~1e-7) past the lnL tolerance
        dfiles.append(f)
        n = {"k": np.linspace(0.05, 6.0, 9), "errs": rng.uniform(5.0, 50.0, 9)}
        nf = str(tmp_path / f"n{i}.npz")
        np.savez(nf, **n)
        nfiles.append(nf)
        noise.append(n)
    up = core_mod.UserParams(BOX_LEN=box_len, HII_DIM=shape[0])
    c = core_mod.CoreLightConeModule(lambda params: state["box"], up)
    lk = lk_mod.Likelihood1DPowerLightcone(nchunks=nchunks, n_psbins=nb, min_k=min_k, max_k=max_k, datafile=dfiles,
                                           noisefile=nfiles, error_on_model=error_on_model, model_uncertainty=0.2)
    chain = chain_mod.build_computation_chain(c, lk, setup=True)
    state["box"] = synth_box(shape, 5)
    model = lk.reduce_data(chain.build_model_data())
    lnl = lk.computeLikelihood(model)
    model_o = O.reduce_data_lightcone(state["box"], box_len, cell, nchunks, nb)
    data_read = [dict(np.load(f)) for f in dfiles]
    lnl_o = O.compute_likelihood(model_o, data_read, noise, min_k, max_k, 0.2, error_on_model)
    assert np.isfinite(lnl_o) and lnl_o < 0 and abs(lnl - lnl_o) <= LNL_RTOL * abs(lnl_o), (lnl, lnl_o)
    # the fused batched path (k_finalize with the noise^2 / error_on_model branches) agrees with the oracle too
    lnls, _ = lk.likelihood_batch([chain.build_model_data(), chain.build_model_data()])
    assert np.all(np.abs(lnls - lnl_o) <= LNL_RTOL * abs(lnl_o))
    return lnl, lnl_o


def check_coeval_staging_slices(mods, monkeypatch, n=10, nz=3, nwalk=4):
    """With several redshifts (data sets) and a host batch staged in several slices, box b must be compared with
    data set b % nsets whatever slice it travels in."""
    lk_mod, core_mod, chain_mod = mods
    up = core_mod.UserParams(BOX_LEN=1.5 * n, HII_DIM=n)

    class Sim:
        def __init__(self):
            self.seed = 0

        def __call__(self, params):
            return [synth_box((n, n, n), 10 * self.seed + z) for z in range(nz)]

    sim = Sim()
    c = core_mod.CoreCoevalModule([7.0, 8.0, 9.0][:nz], sim, up)
    lk = lk_mod.Likelihood1DPowerCoeval(n_psbins=5, simulate=True, min_k=0.3, max_k=4.0)
    chain = chain_mod.build_computation_chain(c, lk)
    ctxs = []
    for w in range(nwalk):
        sim.seed = w + 1
        ctxs.append(chain.build_model_data())
    single = [lk.computeLikelihood(lk.reduce_data(ctx)) for ctx in ctxs]
    monkeypatch.setenv("PS21_STAGE_BYTES", str(2 * n * n * n * 4))  # two cubes per slice
    lnls, _ = lk.likelihood_batch(ctxs)
    assert np.allclose(lnls, single, rtol=1e-9, atol=0) and len(set(np.round(lnls, 6))) == nwalk
