"""Minimal chain / context the likelihood plugs into.

The reference drives its modules through ``cosmoHammer`` (third party, not installable here):
``LikelihoodComputationChain`` (src/py21cmmc/cosmoHammer.py:533-732) creates a ``ChainContext`` per
parameter vector, calls every core, then ``reduce_data`` -> ``store`` -> ``computeLikelihood`` on every
likelihood (``invokeLikelihoodModule``, :652-670).  This module offers the members of those classes
that the PS-likelihood path touches so the plugin API can be exercised without
cosmoHammer, and adds one thing the reference lacks: ``call_batch`` -- evaluate many walkers with a
single batched device call (the adapter SURVEY.md section 8f-1 asks for).
"""
from __future__ import annotations

import numpy as np


class Params:
    """``Params(("name", [guess, min, max, width]), ...)`` (mcmc.py:382, cosmoHammer.py:511-530)."""

    def __init__(self, *pairs):
        self.keys = [k for k, _ in pairs]
        self._vals = [v for _, v in pairs]
        try:
            self.values = np.array(self._vals, dtype=float)
        except (TypeError, ValueError):
            self.values = np.array(self._vals, dtype=object)

    def __getitem__(self, idx):
        return self.values[idx]

    def __getattr__(self, name):
        keys = self.__dict__.get("keys", [])
        if name in keys:
            return self.__dict__["_vals"][keys.index(name)]
        raise AttributeError(name)

    def __len__(self):
        return len(self.keys)

    def items(self):
        yield from zip(self.keys, self._vals)

    def __eq__(self, other):
        if self.__class__.__name__ != other.__class__.__name__ or self.keys != other.keys:
            return False
        return all(np.all(np.asarray(a) == np.asarray(b)) for a, b in zip(self._vals, other._vals))

    __hash__ = object.__hash__


class ChainContext:
    """get / add / contains / remove / getParams / getData, as ``cosmoHammer.ChainContext``."""

    def __init__(self, parent, params):
        self._parent = parent
        self._params = params
        self._store = {}
        self._data = {}

    def add(self, key, value):
        self._store[key] = value

    def get(self, key, default=None):
        return self._store.get(key, default)

    def contains(self, key):
        return key in self._store

    def remove(self, key):
        del self._store[key]

    def getParams(self):
        return self._params

    def getParent(self):
        return self._parent

    def getData(self):
        return self._data


class LikelihoodComputationChain:
    """Cores -> likelihoods orchestration (cosmoHammer.py:533-732, reduced to the PS path)."""

    def __init__(self, params=None):
        self.params = params
        self._setup = False
        self._cores = []
        self._likelihoods = []

    # cosmoHammer accessors
    def getCoreModules(self):
        return self._cores

    def getLikelihoodModules(self):
        return self._likelihoods

    def addCoreModule(self, module):
        self._cores.append(module)
        module._LikelihoodComputationChain = self

    def addLikelihoodModule(self, module):
        self._likelihoods.append(module)
        module._LikelihoodComputationChain = self

    def isValid(self, p):
        if self.params is None:
            return True
        p = np.asarray(p, dtype=float)
        return bool(np.all(p >= self.params[:, 1]) and np.all(p <= self.params[:, 2]))

    def createChainContext(self, p=None):
        if p is None:
            p = {}
        try:
            p = Params(*zip(self.params.keys, p))
        except Exception:
            pass
        return ChainContext(self, p)

    def setup(self):
        if self._setup:
            return
        for m in self._cores + self._likelihoods:
            if hasattr(m, "setup"):
                m.setup()
                m._is_setup = True
        self._setup = True

    def build_model_data(self, p=None):
        ctx = self.createChainContext(p)
        for core in self._cores:
            core.build_model_data(ctx)
        return ctx

    def simulate_mock(self, p=None):
        ctx = self.createChainContext(p)
        for core in self._cores:
            core.simulate_mock(ctx)
        return ctx

    def invokeCoreModule(self, core, ctx):
        if not self._setup:
            self.setup()
        core(ctx)
        core.prepare_storage(ctx, ctx.getData())

    def invokeLikelihoodModule(self, module, ctx):
        if not self._setup:
            self.setup()
        model = module.reduce_data(ctx)
        if hasattr(module, "store"):
            module.store(model, ctx.getData())
        return module.computeLikelihood(model)

    def computeLikelihoods(self, ctx):
        return sum(self.invokeLikelihoodModule(lk, ctx) for lk in self._likelihoods)

    def __call__(self, p):
        if not self.isValid(p):
            return -np.inf, []
        ctx = self.createChainContext(p)
        for core in self._cores:
            self.invokeCoreModule(core, ctx)
        return self.computeLikelihoods(ctx), ctx.getData()

    # ----------------------------------------------------------------- batched walkers (section 8f-1)
    def call_batch(self, plist):
        """Evaluate many parameter vectors: cores per walker, then ONE batched call per likelihood.

        Returns ``[(lnL, blobs), ...]`` in the order of ``plist`` -- the shape ``pool.map(lnpostfn, plist)``
        yields for the reference's emcee sampler (ensemble.py:204-208) -- so an object exposing
        ``map(fn, plist) = chain.call_batch(plist)`` can stand in for the process pool.
        """
        if not self._setup:
            self.setup()
        n = len(plist)
        lnl = np.zeros(n)
        blobs = [dict() for _ in range(n)]
        valid = [i for i, p in enumerate(plist) if self.isValid(p)]
        ctxs = {}
        for i in valid:
            ctx = self.createChainContext(plist[i])
            for core in self._cores:
                self.invokeCoreModule(core, ctx)
            ctxs[i] = ctx
        for lk in self._likelihoods:
            if hasattr(lk, "reduce_data_batch"):
                vals, models = lk.likelihood_batch([ctxs[i] for i in valid])
            else:
                models = [lk.reduce_data(ctxs[i]) for i in valid]
                vals = [lk.computeLikelihood(m) for m in models]
            for j, i in enumerate(valid):
                if hasattr(lk, "store"):
                    lk.store(models[j], ctxs[i].getData())
                lnl[i] += vals[j]
        out = []
        for i in range(n):
            if i in ctxs:
                out.append((float(lnl[i]), ctxs[i].getData()))
            else:
                out.append((-np.inf, []))
        return out


def build_computation_chain(core_modules, likelihood_modules, params=None, setup=True):
    """mcmc.py:21-56."""
    if not isinstance(core_modules, (list, tuple)):
        core_modules = [core_modules]
    if not isinstance(likelihood_modules, (list, tuple)):
        likelihood_modules = [likelihood_modules]
    chain = LikelihoodComputationChain(params)
    for cm in core_modules:
        chain.addCoreModule(cm)
    for lk in likelihood_modules:
        chain.addLikelihoodModule(lk)
    if setup:
        chain.setup()
    return chain
