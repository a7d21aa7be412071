"""Plugin base classes (mirror of src/py21cmmc/core.py:50-223) and box-providing cores.

The reference's ``CoreLightConeModule`` / ``CoreCoevalModule`` (core.py:226-561) run 21cmFAST, which is
outside this package's scope.  The cores here honour the same contract
towards the likelihood -- ``ctx.add("lightcone", lc)`` with ``lc.brightness_temp / n_slices /
cell_size`` (core.py:561, likelihood.py:834-856) or ``ctx.add("brightness_temp", [cube, ...])``
(core.py:436-438) plus ``user_params.BOX_LEN`` -- but take the boxes from a user callable
(a simulator wrapper, a file reader, or a synthetic generator in tests and benchmarks).
"""
from __future__ import annotations

import inspect
import logging

import numpy as np

logger = logging.getLogger("21cmFAST")


class NotSetupError(AttributeError):
    def __init__(self):
        super().__init__("setup() must have been called on the chain to use this method/attribute!")


class NotAChain(AttributeError):
    def __init__(self):
        super().__init__("this Core or Likelihood must be part of a LikelihoodComputationChain "
                         "to enable this method/attribute!")


def _reference_core():
    """``py21cmmc.core`` when the reference is installed next to this package (a drop-in must
    plug into the reference's own chain), else None."""
    try:
        import py21cmmc.core as rc

        return rc if hasattr(rc, "ModuleBase") and hasattr(rc, "CoreBase") else None
    except Exception:
        return None


_REF = _reference_core()


class _ModuleBase:
    """Base of cores and likelihoods -- the plugin contract of core.py:50-138, used when py21cmmc is not importable
    (it needs py21cmfast / cosmoHammer / powerbox); with the reference installed its own class is the base."""

    _extra_defining_attributes = ()
    _ignore_attributes = ()
    required_cores = ()

    def __init__(self):
        self._is_setup = False

    def _check_required_cores(self):
        for rc in self.required_cores:
            if not hasattr(rc, "__len__"):
                rc = (rc,)
            if not any(any(isinstance(m, r) for r in rc) for m in self._cores):
                raise ValueError("%s needs the %s to be loaded." % (self.__class__.__name__, rc.__class__.__name__))

    @property
    def chain(self):
        try:
            return self._LikelihoodComputationChain
        except AttributeError:
            raise NotAChain

    @property
    def parameter_names(self):
        return getattr(self.chain.params, "keys", [])

    def __eq__(self, other):
        if self.__class__.__name__ != other.__class__.__name__:
            return False
        args = []
        for cls in self.__class__.mro():
            args += inspect.getfullargspec(cls.__init__).args
        for arg in tuple(set(args)) + tuple(self._extra_defining_attributes):
            if arg == "self" or arg in self._ignore_attributes:
                continue
            try:
                if hasattr(self, "_" + arg):
                    if getattr(self, "_" + arg) != getattr(other, "_" + arg):
                        return False
                elif hasattr(self, arg):
                    if getattr(self, arg) != getattr(other, arg):
                        return False
                else:
                    logger.warning(f"parameter {arg} not found in instance")
            except ValueError:
                logger.warning(f"parameter {arg} has type which does not allow for comparison")
        return True

    __hash__ = object.__hash__

    @property
    def _cores(self):
        return self.chain.getCoreModules()

    @property
    def _rq_cores(self):
        req = []
        for rc in self.required_cores:
            req.extend(rc if hasattr(rc, "__len__") else (rc,))
        return tuple(core for core in self._cores for r in req if isinstance(core, r))

    @property
    def core_primary(self):
        return self._rq_cores[0] if self._rq_cores else self._cores[0]

    def setup(self):
        self._check_required_cores()


ModuleBase = _REF.ModuleBase if _REF is not None else _ModuleBase


class _CoreBase(ModuleBase):
    """core.py:148-223 (fallback, see _ModuleBase)."""

    def __init__(self, store=None):
        super().__init__()
        self.store = store or {}

    def _check_required_cores(self):
        pass

    def prepare_storage(self, ctx, storage):
        for name, fn in self.store.items():
            storage[name] = fn(ctx)

    def build_model_data(self, ctx):
        pass

    def convert_model_to_mock(self, ctx):
        pass

    def simulate_mock(self, ctx):
        self.build_model_data(ctx)
        self.convert_model_to_mock(ctx)

    def __call__(self, ctx):
        self.build_model_data(ctx)


CoreBase = _REF.CoreBase if _REF is not None else _CoreBase


class UserParams:
    """The two members of ``py21cmfast.UserParams`` the likelihood reads (likelihood.py:645, :856)."""

    def __init__(self, BOX_LEN, HII_DIM):
        self.BOX_LEN = float(BOX_LEN)
        self.HII_DIM = int(HII_DIM)

    def __eq__(self, other):
        return isinstance(other, UserParams) and (self.BOX_LEN, self.HII_DIM) == (other.BOX_LEN, other.HII_DIM)

    def __ne__(self, other):
        return not self.__eq__(other)

    def __repr__(self):
        return f"UserParams(BOX_LEN={self.BOX_LEN}, HII_DIM={self.HII_DIM})"


class LightCone:
    """Duck type of ``py21cmfast.LightCone``: brightness_temp (HII_DIM, HII_DIM, n_slices), cell_size."""

    def __init__(self, brightness_temp, cell_size):
        self.brightness_temp = brightness_temp
        self.cell_size = float(cell_size)

    @property
    def n_slices(self):
        return self.brightness_temp.shape[-1]


class CoreLightConeModule(CoreBase):
    """Puts a lightcone into the context (contract of core.py:531-561).

    ``simulator(params) -> float32 array (HII_DIM, HII_DIM, n_slices)``; ``params`` is the context's
    parameter object (``ctx.getParams()``).
    """

    def __init__(self, simulator, user_params, name="", store=None):
        super().__init__(store)
        self.simulator = simulator
        self.user_params = user_params if isinstance(user_params, UserParams) else UserParams(**user_params)
        self.name = name

    def build_model_data(self, ctx):
        bt = self.simulator(ctx.getParams())
        bt = np.asarray(bt)
        ctx.add("lightcone", LightCone(bt, self.user_params.BOX_LEN / self.user_params.HII_DIM))


class CoreCoevalModule(CoreBase):
    """Puts a list of coeval cubes (one per redshift) into the context (contract of core.py:413-440)."""

    def __init__(self, redshift, simulator, user_params, name="", store=None):
        super().__init__(store)
        self.redshift = list(redshift) if hasattr(redshift, "__len__") else [redshift]
        self.simulator = simulator
        self.user_params = user_params if isinstance(user_params, UserParams) else UserParams(**user_params)
        self.name = name

    def build_model_data(self, ctx):
        boxes = self.simulator(ctx.getParams())
        if len(boxes) != len(self.redshift):
            raise ValueError("simulator must return one cube per redshift")
        ctx.add("brightness_temp", [np.asarray(b) for b in boxes])


# What the likelihoods accept as their ``required_cores`` / ``paired_core``: the box-providing cores of this module
# and, when the reference is importable, its own 21cmFAST-running ones (likelihood.py:340, :695, :666-682, :878-892).
LIGHTCONE_CORES = (CoreLightConeModule,) + ((_REF.CoreLightConeModule,) if _REF is not None and
                                             hasattr(_REF, "CoreLightConeModule") else ())
COEVAL_CORES = (CoreCoevalModule,) + ((_REF.CoreCoevalModule,) if _REF is not None and
                                       hasattr(_REF, "CoreCoevalModule") else ())
