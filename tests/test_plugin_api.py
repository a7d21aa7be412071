"""Plugin-API behaviour mirrored from the reference (src/py21cmmc/likelihood.py:136-266, 342-431, 584-607,
697-732; core.py:92-122): data/noise files, simulate=True, format and length checks, error_on_model, noise
splines, equality, pickling.  Runs on the emulated kernels (CPU)."""
import os
import pickle

import numpy as np
import pytest

from oracle import ps_oracle as O
from tests.helpers import LNL_RTOL, synth_box

SHAPE, BOX_LEN, NCHUNKS, NB = (12, 12, 36), 18.0, 3, 6


@pytest.fixture(scope="module")
def mods():
    from tests.emu.build_emu import build

    lib = build()
    import py21cmmc_b200  # noqa: F401
    from py21cmmc_b200 import _capi, chain, core, likelihood, plan

    _capi._testing_set_library(lib)
    yield likelihood, core, chain
    plan.clear_plans()
    _capi._testing_set_library(None)


class StateSim:
    """Picklable stand-in for the simulator: returns whatever box the shared state dict holds."""

    def __init__(self, state):
        self.state = state

    def __call__(self, params):
        return self.state["box"]


def make(mods, state, **kw):
    lk_mod, core_mod, chain_mod = mods
    up = core_mod.UserParams(BOX_LEN=BOX_LEN, HII_DIM=SHAPE[0])
    c = core_mod.CoreLightConeModule(StateSim(state), up)
    kw.setdefault("nchunks", NCHUNKS)
    kw.setdefault("n_psbins", NB)
    kw.setdefault("min_k", 0.3)
    kw.setdefault("max_k", 4.0)
    lk = lk_mod.Likelihood1DPowerLightcone(**kw)
    return chain_mod.build_computation_chain(c, lk, setup=False), lk


def test_simulate_writes_datafiles_and_rereads_them(mods, tmp_path):
    files = [str(tmp_path / f"data_{i}.npz") for i in range(NCHUNKS)]
    state = {"box": synth_box(SHAPE, 0)}
    chain, lk = make(mods, state, datafile=files, simulate=True)
    chain.setup()
    assert all(os.path.exists(f) for f in files)
    d0 = dict(np.load(files[0]))
    # the reference's keys plus the bin-edge rule the file was simulated under (ignored by the reference's reader)
    assert set(d0) == {"k", "delta", "edge_rule"} and str(d0["edge_rule"]) == "legacy_1.001"
    assert np.array_equal(d0["delta"], lk.data[0]["delta"], equal_nan=True)
    # a second run moves the old files to .bk (likelihood.py:236-247)
    chain2, lk2 = make(mods, state, datafile=files, simulate=True)
    chain2.setup()
    assert os.path.exists(files[0] + ".bk")
    # reading the files back instead of simulating gives the same likelihood
    chain3, lk3 = make(mods, state, datafile=files, simulate=False)
    chain3.setup()
    state["box"] = synth_box(SHAPE, 3)
    a = lk.computeLikelihood(lk.reduce_data(chain.build_model_data()))
    b = lk3.computeLikelihood(lk3.reduce_data(chain3.build_model_data()))
    assert a == b and a < 0


def test_datafile_written_under_another_edge_rule_is_reported(mods, tmp_path, caplog):
    files = [str(tmp_path / f"data_{i}.npz") for i in range(NCHUNKS)]
    state = {"box": synth_box(SHAPE, 0)}
    chain, _ = make(mods, state, datafile=files, simulate=True, edge_rule="inscribed")
    chain.setup()
    chain2, _ = make(mods, state, datafile=files, simulate=False, edge_rule="legacy_1.001")
    with caplog.at_level("WARNING", logger="21cmFAST"):
        chain2.setup()
    assert any("different k grids" in r.message for r in caplog.records)


def test_setup_errors_like_reference(mods, tmp_path):
    state = {"box": synth_box(SHAPE, 0)}
    chain, _ = make(mods, state)  # neither datafile nor simulate: datafile="" -> [""] is looked up, as in the reference
    with pytest.raises(FileNotFoundError):
        chain.setup()
    co = mods[0].Likelihood1DPowerCoeval()  # coeval default datafile=None -> the explicit ValueError
    chc = mods[2].build_computation_chain(
        mods[1].CoreCoevalModule([7.0], lambda p: [synth_box((8, 8, 8), 0)], mods[1].UserParams(12.0, 8)), co,
        setup=False)
    with pytest.raises(ValueError):
        chc.setup()
    chain, _ = make(mods, state, datafile=[str(tmp_path / "missing.npz")] * NCHUNKS)
    with pytest.raises(FileNotFoundError):
        chain.setup()
    one = str(tmp_path / "one.npz")
    np.savez(one, k=np.linspace(0.1, 1, 5), delta=np.ones(5))
    chain, _ = make(mods, state, datafile=[one])  # 1 file for 3 chunks
    with pytest.raises(ValueError):
        chain.setup()
    bad = str(tmp_path / "bad.npz")
    np.savez(bad, ks=np.linspace(0.1, 1, 5), delta=np.ones(5))  # wrong key
    chain, _ = make(mods, state, datafile=[bad] * NCHUNKS)
    with pytest.raises(ValueError):
        chain.setup()
    with pytest.raises(ValueError):
        mods[0].Likelihood1DPowerCoeval(datafile=["a", "b"], noisefile=["n"])
    # a likelihood outside a chain has no cores
    with pytest.raises(AttributeError):
        mods[0].Likelihood1DPowerLightcone(simulate=True).setup()
    # required core missing
    lk = mods[0].Likelihood1DPowerCoeval(simulate=True)
    ch = mods[2].LikelihoodComputationChain(None)
    ch.addCoreModule(mods[1].CoreLightConeModule(lambda p: state["box"], mods[1].UserParams(BOX_LEN, SHAPE[0])))
    ch.addLikelihoodModule(lk)
    with pytest.raises(ValueError):
        ch.setup()


@pytest.mark.parametrize("error_on_model", [True, False])
def test_noise_and_error_on_model_match_oracle(mods, tmp_path, error_on_model):
    from tests import parity_cases as PC

    PC.check_noise_and_error_on_model(mods, tmp_path, error_on_model, SHAPE, BOX_LEN, NCHUNKS, NB)


def test_equality_and_pickle(mods):
    state = {"box": synth_box(SHAPE, 0)}
    chain, lk = make(mods, state, simulate=True)
    chain.setup()
    other = mods[0].Likelihood1DPowerLightcone(nchunks=NCHUNKS, n_psbins=NB, min_k=0.3, max_k=4.0, simulate=False)
    assert lk == other  # `simulate` is ignored in equality (likelihood.py:134)
    assert lk != mods[0].Likelihood1DPowerLightcone(nchunks=NCHUNKS, n_psbins=NB + 1, min_k=0.3, max_k=4.0)
    assert lk != mods[0].Likelihood1DPowerCoeval(n_psbins=NB)
    lnl = lk.computeLikelihood(lk.reduce_data(chain.build_model_data()))
    # the chain (cores + likelihood, after a device evaluation) pickles: no device handle lives on the instances
    lk2 = pickle.loads(pickle.dumps(lk))
    assert lk2 == lk and lk2.nchunks == NCHUNKS
    assert [np.array_equal(a["delta"], b["delta"], equal_nan=True) for a, b in zip(lk.data, lk2.data)]
    assert lnl == 0.0  # own data


def test_fiducial_model_and_core_params_exposed(mods):
    state = {"box": synth_box(SHAPE, 0)}
    chain, lk = make(mods, state, simulate=True)
    chain.setup()
    assert lk.user_params.BOX_LEN == BOX_LEN
    fid = lk.get_fiducial_model()
    assert len(fid) == NCHUNKS and set(fid[0]) == {"k", "delta"}
    assert lk.paired_core is chain.getCoreModules()[0]


def test_blob_handoff_matches_reference_storage_layout(mods, tmp_path):
    """call_batch -> structured rows shaped like HDFStorage's `blobs` dataset (cosmoHammer.py:227-242, 281-288)."""
    from py21cmmc_b200.chain import Params
    from py21cmmc_b200.walkers import BatchedWalkerPool, BlobChain, blob_dtype

    lk_mod, core_mod, chain_mod = mods
    up = core_mod.UserParams(BOX_LEN=BOX_LEN, HII_DIM=SHAPE[0])

    class SeedSim:
        def __call__(self, params):
            return synth_box(SHAPE, int(round(params.seed)) if len(params) else 0)

    c = core_mod.CoreLightConeModule(SeedSim(), up)
    lk = lk_mod.Likelihood1DPowerLightcone(nchunks=NCHUNKS, n_psbins=NB, simulate=True, min_k=0.3, max_k=4.0)
    chain = chain_mod.build_computation_chain(c, lk, params=Params(("seed", [0.0, 0.0, 50.0, 1.0])))
    pool = BlobChain()
    mapper = BatchedWalkerPool(chain)
    for it in range(2):
        res = mapper.map(None, [[1.0 + it], [2.0 + it], [99.0], [3.0 + it]])  # 99 is outside the prior -> -inf, no blobs
        assert res[2][0] == -np.inf and not res[2][1]
        pool.save_step([r[1] for r in res])
    arr = pool.array()
    assert arr.shape == (2, 4) and set(arr.dtype.names) == {f"{k}_{i}" for k in ("k", "delta") for i in range(NCHUNKS)}
    assert arr.dtype == blob_dtype(res[0][1]) and arr["delta_1"].shape == (2, 4, NB)
    assert np.all(np.isnan(arr["delta_0"][:, 2])) and np.array_equal(arr["k_0"][0, 0], res[0][1]["k_0"])
    # walker 1 of iteration 0 and walker 0 of iteration 1 saw the same lightcone (seed 2)
    assert np.array_equal(arr["delta_2"][0, 1], arr["delta_2"][1, 0], equal_nan=True)
    pool.save(str(tmp_path / "blobs.npy"))
    back = np.load(str(tmp_path / "blobs.npy"))
    assert back.dtype == arr.dtype and np.array_equal(back["delta_0"], arr["delta_0"], equal_nan=True)


def test_coeval_data_sets_follow_boxes_across_staging_slices(mods, monkeypatch):
    """Regression: box b is compared with data set b % nsets whatever staging slice it travels in."""
    from tests import parity_cases as PC

    PC.check_coeval_staging_slices(mods, monkeypatch)


def test_call_batch_composes_with_upper_limit_and_windowed_likelihoods(mods, tmp_path):
    """f-1 (walker batches) and f-2 (HERA-style consumers) must compose: chain.call_batch on the
    Upper / Windowed classes takes their own likelihood, not the Gaussian chi^2 of the parent class."""
    from tests import parity_cases as PC

    lk_mod, core_mod, chain_mod = mods
    shape, box_len, nchunks, nb = (16, 16, 48), 24.0, 3, 8
    up = core_mod.UserParams(BOX_LEN=box_len, HII_DIM=shape[0])

    class SeedSim:
        def __call__(self, params):
            return synth_box(shape, int(round(params.seed)))

    from py21cmmc_b200.chain import Params

    probe = lk_mod.Likelihood1DPowerLightcone(nchunks=nchunks, n_psbins=nb, use_data=False)
    ch0 = chain_mod.build_computation_chain(core_mod.CoreLightConeModule(lambda p: synth_box(shape, 1), up), probe)
    kmodel = probe.reduce_data(ch0.build_model_data())[0]["k"]
    PC.synthetic_hera_file(str(tmp_path / "hera.npz"), kmodel)
    for cls in (lk_mod.Likelihood1DPowerLightconeUpper, lk_mod.Likelihood1DPowerLightconeWindowed):
        lk = cls(datafile=str(tmp_path / "hera.npz"), nchunks=nchunks, n_psbins=nb, band_chunks=[2, 0])
        chain = chain_mod.build_computation_chain(core_mod.CoreLightConeModule(SeedSim(), up), lk,
                                                  params=Params(("seed", [0.0, 0.0, 50.0, 1.0])))
        plist = [[1.0], [2.0], [77.0], [3.0]]   # 77 is outside the prior
        batch = chain.call_batch(plist)
        assert batch[2][0] == -np.inf and not batch[2][1]
        for p, (lnl, blobs) in zip(plist, batch):
            if p[0] > 50:
                continue
            one, blobs1 = chain(p)
            assert np.isfinite(one) and abs(lnl - one) <= 1e-12 * abs(one)
            assert set(blobs) == set(blobs1) and np.array_equal(blobs["delta_1"], blobs1["delta_1"], equal_nan=True)
        # reduce_data_batch == [reduce_data(ctx)]
        ctxs = [chain.build_model_data(p) for p in ([1.0], [3.0])]
        many = lk.reduce_data_batch(ctxs)
        for ctx, m in zip(ctxs, many):
            assert np.array_equal(m[0]["delta"], lk.reduce_data(ctx)[0]["delta"])


def test_host_pointer_path_equals_array_path_and_accepts_pinned_buffers(mods):
    """ps21_lnl_host_ptrs / ps21_power_host_ptrs: per-walker buffers by pointer (no gather on the host) give the very
    results of the contiguous-array entry points; page-locked helpers round-trip."""
    from py21cmmc_b200 import _capi
    from py21cmmc_b200.plan import PowerPlan, pin_array, pinned_empty

    shape = (10, 10, 24)
    pl = PowerPlan(shape, [(0, 12), (12, 24)], [(15.0, 15.0, 18.0)] * 2, 5, library=_capi.LIB_PATH)
    boxes = [synth_box(shape, s) for s in range(5)]
    pinned = pinned_empty(shape)
    pinned[...] = boxes[1]
    boxes[1] = pinned
    unpin = pin_array(boxes[2])
    strided = np.asfortranarray(boxes[3])          # not C-contiguous: converted, the others go zero copy
    f64 = boxes[4].astype(np.float64)              # another dtype: converted
    as_list = [boxes[0], boxes[1], boxes[2], strided, f64]
    stack = np.stack([np.asarray(b, dtype=np.float32) for b in boxes])
    a = pl.power_host(stack)
    b = pl.power_host(as_list)
    assert np.array_equal(a, b, equal_nan=True)
    data = a[0].reshape(1, 2, 5) * 1.3
    pl.set_data(np.nan_to_num(data), None, np.isfinite(data), 0.15, True)
    la, da, _ = pl.lnl_host(stack, want_delta=True)
    lb, db, pb = pl.lnl_host(as_list, want_delta=True, want_power=True)
    assert np.array_equal(la, lb) and np.array_equal(da, db, equal_nan=True) and np.array_equal(pb, a, equal_nan=True)
    with pytest.raises(ValueError):
        pl.power_host([np.zeros((10, 10, 23), np.float32)])
    unpin()
    # giving the workspace back keeps the plan usable
    pl.release_workspace()
    assert pl.info()["workspace_bytes"] == 0
    assert np.array_equal(pl.power_host(as_list), a, equal_nan=True)
    pl.close()


def test_two_likelihoods_sharing_a_cached_plan_keep_their_own_data(mods):
    """plans are cached per geometry and shared; the data tables are bound to the call."""
    lk_mod, core_mod, chain_mod = mods
    state = {"box": synth_box(SHAPE, 0)}
    chain_a, lk_a = make(mods, state, simulate=True, min_k=0.3, max_k=4.0)
    chain_a.setup()
    chain_b, lk_b = make(mods, state, simulate=True, min_k=0.8, max_k=3.0, model_uncertainty=0.3)
    chain_b.setup()
    state["box"] = synth_box(SHAPE, 4)
    ctx = chain_a.build_model_data()
    assert lk_a._geometry(ctx)[1] is lk_b._geometry(ctx)[1]      # one plan, two likelihoods
    want_a = lk_a.computeLikelihood(lk_a.reduce_data(ctx))
    want_b = lk_b.computeLikelihood(lk_b.reduce_data(ctx))
    assert want_a != want_b
    plan = lk_a._geometry(ctx)[1]
    for _ in range(3):   # interleaved: every call re-binds its own tables; repeated calls skip the upload
        assert abs(lk_a.likelihood_batch([ctx])[0][0] - want_a) <= LNL_RTOL * abs(want_a)
        key_a = plan._data_key
        assert abs(lk_a.likelihood_batch([ctx, ctx])[0][1] - want_a) <= LNL_RTOL * abs(want_a)
        assert plan._data_key is key_a
        assert abs(lk_b.likelihood_batch([ctx])[0][0] - want_b) <= LNL_RTOL * abs(want_b)
    pickle.loads(pickle.dumps(lk_a))   # the per-process table cache is not pickled


def test_plan_cache_is_bounded(mods, monkeypatch):
    from py21cmmc_b200 import plan as plan_mod

    monkeypatch.setattr(plan_mod, "_PLAN_CACHE_MAX", 2)
    plan_mod.clear_plans()
    plans = [plan_mod.get_plan((8, 8, 8 + 2 * i), [(0, 8 + 2 * i)], [(12.0, 12.0, 12.0 + 3 * i)], 4) for i in range(4)]
    assert len(plan_mod._PLANS) == 2
    first = plans[0]
    box = synth_box(first.shape, 0)
    want = first.power_host(box)            # an evicted plan still works for whoever holds it
    assert np.isfinite(want).any()
    assert plan_mod.get_plan((8, 8, 14), [(0, 14)], [(12.0, 12.0, 21.0)], 4) is plans[3]
    plan_mod.clear_plans()


def _seed_chain(mods, prior_max=50.0):
    from py21cmmc_b200.chain import Params

    lk_mod, core_mod, chain_mod = mods
    up = core_mod.UserParams(BOX_LEN=BOX_LEN, HII_DIM=SHAPE[0])

    class SeedSim:
        def __call__(self, params):
            return synth_box(SHAPE, int(round(params.seed)) if len(params) else 0)

    c = core_mod.CoreLightConeModule(SeedSim(), up)
    lk = lk_mod.Likelihood1DPowerLightcone(nchunks=NCHUNKS, n_psbins=NB, simulate=True, min_k=0.3, max_k=4.0)
    return chain_mod.build_computation_chain(c, lk, params=Params(("seed", [0.0, 0.0, prior_max, 1.0])))


def test_vectorised_sampler_entry_points_and_half_ensembles(mods):
    """zeus ``vectorize=True`` / UltraNest ``vectorized=True`` hand all points of a step to the log-probability at
    once (mcmc.py:495-511, 517-545); emcee's stretch move evaluates half the ensemble per call (ensemble.py:166-170)."""
    from py21cmmc_b200.walkers import BatchedWalkerPool, ultranest_prior_transform, vectorized_log_prob

    chain = _seed_chain(mods)
    f = vectorized_log_prob(chain)
    P = np.array([[1.0], [2.0], [77.0], [3.0], [4.0]])      # 77: outside the prior -> -inf, cores not run
    lnl = f(P)
    assert lnl.shape == (5,) and lnl[2] == -np.inf and np.all(np.isfinite(np.delete(lnl, 2)))
    for i in (0, 1, 3, 4):
        one = chain(list(P[i]))[0]                            # chi^2 summed per chunk there, in one tree here
        assert abs(lnl[i] - one) <= 1e-12 * abs(one)
    assert f(P[1]) == lnl[1]                                  # a single point works too
    # zeus with blobs_dtype: (lnL, structured rows)
    fb = vectorized_log_prob(chain, with_blobs=True)
    l2, blobs = fb(P)
    assert np.array_equal(l2, lnl) and blobs.shape == (5,) and np.all(np.isnan(blobs["delta_0"][2]))
    # (Delta^2 = P k^3 / 2 pi^2 is formed on the device in the batch, by NumPy in the single call: last bits may differ)
    assert np.allclose(blobs["delta_1"][3], chain([3.0])[1]["delta_1"], rtol=1e-14, atol=0, equal_nan=True)
    # half-ensemble batches of the stretch move: 5 walkers -> halves of 2 and 3 through the pool interface
    pool = BatchedWalkerPool(chain)
    halfk = len(P) // 2
    first = pool.map(None, [list(p) for p in P[:halfk]])
    second = pool.map(None, [list(p) for p in P[halfk:]])
    assert [r[0] for r in first + second] == list(lnl)
    # UltraNest prior transform (mcmc.py:533-543), vectorised
    t = ultranest_prior_transform(chain.params)(np.array([[0.0], [0.5], [1.0]]))
    assert np.array_equal(t, np.array([[0.0], [25.0], [50.0]]))
    assert np.all(np.isfinite(vectorized_log_prob(chain)(t[1:2])))


def test_hdf_storage_layout_contract(mods, tmp_path):
    """The dataset layout of the reference's HDFStorage (cosmoHammer.py:60-105, 208-295); a real file when h5py is
    importable, the shape / dtype contract otherwise."""
    from py21cmmc_b200.walkers import HDFBlobStorage, blob_dtype, blobs_to_structured

    chain = _seed_chain(mods)
    res = chain.call_batch([[1.0], [2.0], [3.0]])
    blobs = [r[1] for r in res]
    lay = HDFBlobStorage.layout(3, 1, blobs[0])
    assert lay["chain"] == ((0, 3, 1), (None, 3, 1), np.dtype(np.float64))
    assert lay["accepted"][2] == np.dtype(int) and lay["trial_log_prob"][1] == (None, 3)
    assert lay["blobs"][2] == blob_dtype(blobs[0]) and lay["blobs"][2]["delta_0"].shape == (NB,)
    rows = blobs_to_structured(blobs, lay["blobs"][2])
    assert rows.shape == (3,) and np.array_equal(rows["k_2"][1], blobs[1]["k_2"])
    try:
        import h5py
    except ImportError:
        with pytest.raises(ImportError):
            HDFBlobStorage(str(tmp_path / "x.h5"))
        return
    st = HDFBlobStorage(str(tmp_path / "x.h5"))
    st.reset(3, chain.params)
    coords = np.array([[1.0], [2.0], [3.0]])
    for it in range(2):
        st.grow(1, blobs[0])
        st.save_step(coords, [r[0] for r in res], blobs, coords, [r[0] for r in res], [True] * 3)
    with h5py.File(str(tmp_path / "x.h5")) as f:
        g = f["sample_0"]
        assert g.attrs["iteration"] == 2 and g["blobs"].shape == (2, 3) and g["chain"].shape == (2, 3, 1)
        assert np.array_equal(g["blobs"]["delta_0"][1, 2], blobs[2]["delta_0"], equal_nan=True)


def test_reference_core_classes_are_used_when_py21cmmc_imports(monkeypatch):
    """Drop-in boundary: with the reference importable, ModuleBase / CoreBase are ITS classes and its
    CoreLightConeModule / CoreCoevalModule satisfy required_cores / paired_core (likelihood.py:340, :695)."""
    import importlib.util
    import sys
    import types

    import py21cmmc_b200.core as here   # imported without the reference: the local fallback classes serve

    assert here.LIGHTCONE_CORES == (here.CoreLightConeModule,) and here.ModuleBase is here._ModuleBase
    ref = types.ModuleType("py21cmmc.core")

    class RefModuleBase:
        required_cores = ()

        def __init__(self):
            self._is_setup = False

    class RefCoreBase(RefModuleBase):
        def __init__(self, store=None):
            super().__init__()
            self.store = store or {}

    class RefLightCone(RefCoreBase):
        pass

    class RefCoeval(RefCoreBase):
        pass

    ref.ModuleBase, ref.CoreBase = RefModuleBase, RefCoreBase
    ref.CoreLightConeModule, ref.CoreCoevalModule = RefLightCone, RefCoeval
    pkg = types.ModuleType("py21cmmc")
    pkg.core = ref
    monkeypatch.setitem(sys.modules, "py21cmmc", pkg)
    monkeypatch.setitem(sys.modules, "py21cmmc.core", ref)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("ps21_core_probe", os.path.join(root, "21cmmc_b200", "core.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.ModuleBase is RefModuleBase and mod.CoreBase is RefCoreBase
    assert issubclass(mod.CoreLightConeModule, RefCoreBase)
    assert mod.LIGHTCONE_CORES == (mod.CoreLightConeModule, RefLightCone)
    assert mod.COEVAL_CORES == (mod.CoreCoevalModule, RefCoeval)


HERA_FILE = "/root/reference/src/py21cmmc/data/HERA_H1C_IDR3.npz"


@pytest.mark.skipif(not os.path.exists(HERA_FILE), reason="the reference tree (and its HERA data file) is not on this box")
def test_upper_limit_likelihood_on_the_reference_hera_file(mods, monkeypatch):
    """f-2 on the reference's real data file (src/py21cmmc/data/HERA_H1C_IDR3.npz: band8 (5,16,3), band10 (5,19,3),
    window functions 32 / 37 wide, z_bands (2,)) through ``from_builtin_data``, against the oracle's restatement of
    likelihood.py:2249-2335."""
    lk_mod, core_mod, chain_mod = mods
    monkeypatch.setenv("PS21_DATA_DIR", os.path.dirname(HERA_FILE))
    shape, box_len, nchunks, nb = (24, 24, 96), 240.0, 2, 14     # 10 Mpc cells: k reaches the HERA window (0.18-1.7 /Mpc)
    up = core_mod.UserParams(BOX_LEN=box_len, HII_DIM=shape[0])
    state = {"box": synth_box(shape, 3)}
    lk = lk_mod.Likelihood1DPowerLightconeUpper.from_builtin_data("HERA_H1C_IDR3", nchunks=nchunks, n_psbins=nb,
                                                                   band_chunks=[1, 0])
    chain = chain_mod.build_computation_chain(core_mod.CoreLightConeModule(lambda p: state["box"], up), lk)
    assert lk.datafile == [HERA_FILE] and list(np.round(lk.redshifts, 2)) == [10.37, 7.93] and lk.k_len == 37
    model = lk.reduce_data(chain.build_model_data())
    assert model[0]["delta"].shape == (1, 2, 37)
    hera = dict(np.load(HERA_FILE, allow_pickle=True))
    rng = np.random.default_rng(2)
    # many models spanning "far below every limit" to "far above": the erf saturates on both sides
    delta = np.abs(model[0]["delta"]) * rng.uniform(1e-2, 1e5, (12, 1, 1)) + rng.uniform(0.0, 50.0, (12, 2, 37))
    got = lk.computeLikelihood([{"k": lk.k, "delta": delta}])
    want = O.upper_limit_likelihood(hera, hera["z_bands"], delta)
    assert np.all(np.isfinite(want)) and np.ptp(want) > 10     # the models really differ in likelihood
    assert np.allclose(got, want, rtol=1e-9, atol=1e-9)
    err = np.abs(delta[0]) * 0.3
    got_e = lk.computeLikelihood([{"k": lk.k, "delta": delta, "delta_err": err}])
    want_e = O.upper_limit_likelihood(hera, hera["z_bands"], delta, err)
    assert np.allclose(got_e, want_e, rtol=1e-9, atol=1e-9) and not np.allclose(got_e, got)
    # and through the batched chain
    lnl, _ = chain([])
    lnl_b = chain.call_batch([[], []])
    assert np.isfinite(lnl) and lnl_b[0][0] == lnl_b[1][0] and abs(lnl_b[0][0] - lnl) <= 1e-12 * abs(lnl)
