"""Power-spectrum likelihoods with the plugin API of src/py21cmmc/likelihood.py, computed on the GPU.

``Likelihood1DPowerCoeval`` / ``Likelihood1DPowerLightcone`` keep the constructor options, ``setup``,
``reduce_data``, ``computeLikelihood``, ``store`` and the ``compute_power`` staticmethod of the reference
(likelihood.py:269-892).  Where the reference calls ``powerbox.get_power`` per chunk (:480-488,
:782-790) these classes call libps21.so through ctypes; there is no CPU fallback.  Two additions:

* ``reduce_data_batch`` / ``likelihood_batch``: many walkers in one device call (batched FFT +
  binning + chi^2, only scalars and Delta^2 records come back);
* ``device`` / ``edge_rule`` / ``kernel_path`` constructor options (they take part in ``__eq__`` like every
  other init argument, core.py:92-122).

Instances hold no device handle (plans live in a per-process cache), so they pickle into pool
workers and YAML-dump like the reference's.
"""
from __future__ import annotations

import logging
from functools import cached_property
from io import IOBase
from os import path, rename
from pathlib import Path

import numpy as np
from scipy.interpolate import InterpolatedUnivariateSpline

from . import _capi, binning, core
from .plan import get_plan

logger = logging.getLogger("21cmFAST")
np.seterr(invalid="ignore", divide="ignore")  # likelihood.py:22: empty bins are NaN, silently

_DEFAULTS = {"device": None, "edge_rule": "auto", "kernel_path": "auto"}
_AUTO_RULE = []   # the probed rule, once per process


def resolve_edge_rule(rule=None):
    """``"auto"`` / ``None`` -> the top-edge rule of the powerbox installed next to this package.

    The rule changed between powerbox generations and the reference pins only
    ``powerbox>=0.5.7``, so the installed package is probed once per process
    (``binning.detect_edge_rule``).  Without powerbox the legacy rule is used and a warning is logged: datafiles
    written by a reference install with a newer powerbox then sit on a different k grid."""
    if rule is None or rule == "auto":
        rule = _DEFAULTS["edge_rule"]   # an explicit process default (set_defaults) wins over probing
    if rule != "auto":
        if rule not in binning.EDGE_RULES:
            raise ValueError(f"edge_rule must be 'auto' or one of {binning.EDGE_RULES}")
        return rule
    if not _AUTO_RULE:
        found = binning.detect_edge_rule()
        if found is None:
            logger.warning("powerbox is not importable: cannot probe its bin-edge rule, using 'legacy_1.001' "
                           "(pass edge_rule='inscribed' to match powerbox >= 0.7)")
            found = "legacy_1.001"
        _AUTO_RULE.append(found)
    return _AUTO_RULE[0]


def set_defaults(device=None, edge_rule=None, kernel_path=None):
    """Defaults used by the ``compute_power`` staticmethods (which have no instance to ask)."""
    if device is not None:
        _DEFAULTS["device"] = int(device)
    if edge_rule is not None:
        if edge_rule != "auto" and edge_rule not in binning.EDGE_RULES:
            raise ValueError(f"edge_rule must be 'auto' or one of {binning.EDGE_RULES}")
        _DEFAULTS["edge_rule"] = edge_rule
    if kernel_path is not None:
        _DEFAULTS["kernel_path"] = kernel_path


def builtin_data_dir():
    """Directory of the built-in data files (``from_builtin_data``, likelihood.py:702-707, :2192-2197): PS21_DATA_DIR,
    else the reference's own ``py21cmmc/data`` when it is installed (this package ships no observational data), else
    ``21cmmc_b200/data``."""
    import importlib.util
    import os

    if os.environ.get("PS21_DATA_DIR"):
        return os.environ["PS21_DATA_DIR"]
    try:
        spec = importlib.util.find_spec("py21cmmc")
        if spec is not None and spec.submodule_search_locations:
            cand = path.join(list(spec.submodule_search_locations)[0], "data")
            if path.isdir(cand):
                return cand
    except Exception:
        pass
    return path.join(path.dirname(__file__), "data")


def _device(dev=None):
    if dev is not None:
        return int(dev)
    if _DEFAULTS["device"] is not None:
        return _DEFAULTS["device"]
    try:
        import torch

        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:
        pass
    return 0


def _is_cuda_tensor(x):
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


def _power_records(plan, boxes):
    """(batch, out_per_box) float64 power for a host array, a CUDA tensor, or a list of per-walker boxes of either
    kind (host boxes go to the library by pointer: no gather on the host)."""
    if isinstance(boxes, (list, tuple)) and boxes and _is_cuda_tensor(boxes[0]):
        import torch

        boxes = torch.stack([b.contiguous() for b in boxes])
    if _is_cuda_tensor(boxes):
        return plan.power_device(boxes.contiguous()).cpu().numpy()
    return plan.power_host(boxes)


class LikelihoodBase(core.ModuleBase):
    """likelihood.py:40-108."""

    def computeLikelihood(self, model):
        raise NotImplementedError("The Base likelihood should never be used directly!")

    def reduce_data(self, ctx):
        raise NotImplementedError("The Base likelihood should never be used directly!")

    def _expose_core_parameters(self):
        for m in self._cores:
            for k in ("user_params", "flag_options", "cosmo_params", "astro_params"):
                if hasattr(m, k):
                    if hasattr(self, k) and getattr(self, k) != getattr(m, k):
                        raise ValueError("Setup has detected incompatible input parameter dicts in "
                                         f"specified cores: {k}")
                    setattr(self, k, getattr(m, k))

    def setup(self):
        super().setup()
        self._expose_core_parameters()

    def get_fiducial_model(self):
        return self.reduce_data(self.chain.build_model_data())


class LikelihoodBaseFile(LikelihoodBase):
    """Data / noise read from (or, with ``simulate=True``, written to) ``.npz`` files (likelihood.py:111-266)."""

    _ignore_attributes = LikelihoodBase._ignore_attributes + ("simulate",)

    def __init__(self, datafile=None, noisefile=None, simulate=False, use_data=True):
        super().__init__()
        self.datafile = [datafile] if isinstance(datafile, (str, IOBase)) else datafile
        self.noisefile = [noisefile] if isinstance(noisefile, (str, IOBase)) else noisefile
        self._use_data = use_data
        self._simulate = simulate
        self.data = None
        self.noise = None

    def setup(self):
        super().setup()
        if not self._use_data:
            return
        if not self._simulate and not self.datafile:
            raise ValueError("Either an existing datafile has to be specified, or simulate set to True.")
        simctx = self.chain.simulate_mock() if self._simulate else None
        self.data = self.reduce_data(simctx) if self._simulate else self._read_data()
        can_define = hasattr(self, "define_noise") and self._simulate
        if can_define or self.noisefile:
            self.noise = self.define_noise(simctx, self.data) if can_define else self._read_noise()
        if self.datafile and self._simulate:
            self._write(self.datafile, self.data, "data")
        if self.noisefile and can_define:
            self._write(self.noisefile, self.noise, "noise")

    @staticmethod
    def _load_all(files, kind, hint=""):
        out = []
        for fl in files:
            if not path.exists(fl):
                raise FileNotFoundError(f"Could not find {kind}file: {fl}. {hint}")
            out.append(dict(np.load(fl, allow_pickle=True)))
        return out

    def _read_data(self):
        return self._load_all(self.datafile, "data", "If you meant to simulate data, set simulate=True.")

    def _read_noise(self):
        return self._load_all(self.noisefile, "noise") if self.noisefile else None

    def _file_meta(self):
        """Extra keys written next to the reference's ones (ignored by the reference's reader)."""
        return {}

    def _write(self, files, records, kind):
        for fl, d in zip(files, records):
            if not fl:  # the lightcone class defaults to datafile="": nothing to write (the reference would
                continue  # create a file called ".npz" in the working directory here)
            if path.exists(fl):
                logger.warning(f"File {fl} already exists. Moving previous version to {fl}.bk")
                rename(fl, fl + ".bk")
            np.savez(fl, **d, **self._file_meta())
            logger.info(f"Saved {kind} file: {fl}")

    def _write_data(self):
        self._write(self.datafile, self.data, "data")

    def _write_noise(self):
        self._write(self.noisefile, self.noise, "noise")


class Likelihood1DPowerCoeval(LikelihoodBaseFile):
    """Gaussian likelihood on the spherically averaged Delta^2(k) of coeval cubes (likelihood.py:269-682)."""

    required_cores = (core.COEVAL_CORES,)
    _variant = "coeval"

    def __init__(self, n_psbins=None, min_k=0.1, max_k=1.0, logk=True, model_uncertainty=0.15,
                 error_on_model=True, ignore_kperp_zero=True, ignore_kpar_zero=False, ignore_k_zero=False,
                 name="", device=None, edge_rule="auto", kernel_path="auto", *args, **kwargs):
        super().__init__(*args, **kwargs)
        if self.noisefile and self.datafile and len(self.datafile) != len(self.noisefile):
            raise ValueError("If noisefile or datafile are provided, they should have the same number "
                             "of files (one for each coeval box)")
        if edge_rule != "auto" and edge_rule not in binning.EDGE_RULES:
            raise ValueError(f"edge_rule must be 'auto' or one of {binning.EDGE_RULES}")
        self.n_psbins = n_psbins
        self.min_k = min_k
        self.max_k = max_k
        self.logk = logk
        self.error_on_model = error_on_model
        self.model_uncertainty = model_uncertainty
        self.ignore_k_zero = ignore_k_zero
        self.ignore_kperp_zero = ignore_kperp_zero
        self.ignore_kpar_zero = ignore_kpar_zero
        self.name = name
        self.device = device
        self.edge_rule = edge_rule
        self.kernel_path = kernel_path

    # ------------------------------------------------------------------ format checks / setup
    def _check_data_format(self):
        for i, d in enumerate(self.data):
            if "k" not in d or "delta" not in d:
                raise ValueError(f"datafile #{i + 1} of {len(self.data)} has the wrong format.")

    def _check_noise_format(self):
        for i, n in enumerate(self.noise):
            if "k" not in n or "errs" not in n:
                raise ValueError(f"noisefile #{i + 1} of {len(self.noise)} has the wrong format")

    def _n_expected(self):
        return len(self.redshift)

    def _file_meta(self):
        return {"edge_rule": np.array(resolve_edge_rule(self.edge_rule))}

    def _check_edge_rule(self):
        """A datafile simulated under another bin-edge rule sits on a different k grid: say so."""
        mine = resolve_edge_rule(self.edge_rule)
        for i, d in enumerate(self.data or []):
            theirs = d.get("edge_rule") if hasattr(d, "get") else None
            if theirs is not None and str(theirs) != mine:
                logger.warning(f"datafile #{i + 1} was written with bin-edge rule {theirs}, this likelihood uses "
                               f"{mine}: the model and the data are binned on different k grids")

    def setup(self):
        super().setup()
        if not self._use_data:
            return
        what = "redshift" if self._variant == "coeval" else "chunk"
        if len(self.data) != self._n_expected():
            raise ValueError(f"There needs to be one dataset (datafile) for each {what}!")
        if self.noise and len(self.noise) != self._n_expected():
            raise ValueError(f"There needs to be one dataset (noisefile) for each {what}!")
        self._check_data_format()
        self._check_edge_rule()
        if self.noise:
            self._check_noise_format()

    @cached_property
    def data_spline(self):
        return [InterpolatedUnivariateSpline(d["k"], d["delta"], k=1) for d in self.data]

    @cached_property
    def noise_spline(self):
        if not self.noise:
            return None
        return [InterpolatedUnivariateSpline(n["k"], n["errs"], k=1) for n in self.noise]

    @property
    def redshift(self):
        return self.core_primary.redshift

    @cached_property
    def paired_core(self):
        paired = [c for c in self._cores if isinstance(c, core.COEVAL_CORES)]
        if len(paired) > 1:
            raise ValueError("You've got more than one CoreCoevalModule -- they will overwrite each other!")
        return paired[0]

    # ------------------------------------------------------------------ the hot path
    @staticmethod
    def _plan(shape, bounds, lengths, n_psbins, log_bins, ikperp, ikpar, ik0, device=None, edge_rule=None,
              kernel_path=None, mode="spherical"):
        if n_psbins is None:
            cyl = mode == "cylindrical"
            nbs = {binning.default_nbins((shape[0], shape[1], b - a), 2 if cyl else 3) for a, b in bounds}
            if len(nbs) != 1:
                raise NotImplementedError("n_psbins=None gives different default bin counts for chunks of "
                                          "different length; pass n_psbins explicitly")
            n_psbins = nbs.pop()
        return get_plan(shape, bounds, lengths, int(n_psbins), log_bins, ikperp, ikpar, ik0,
                        resolve_edge_rule(edge_rule), mode, _device(device),
                        kernel_path or _DEFAULTS["kernel_path"])

    @staticmethod
    def compute_power(brightness_temp, L, n_psbins, log_bins=True, ignore_kperp_zero=True,
                      ignore_kpar_zero=False, ignore_k_zero=False):
        """Spherically averaged P(k) of one coeval cube: ``[power, k_centres]`` (likelihood.py:433-498)."""
        shape = tuple(brightness_temp.shape)
        if len(set(shape)) != 1:
            raise ValueError("coeval cubes must be cubic (the reference indexes all axes with shape[0]//2)")
        plan = Likelihood1DPowerCoeval._plan(shape, [(0, shape[2])], [(L, L, L)], n_psbins, log_bins,
                                             ignore_kperp_zero, ignore_kpar_zero, ignore_k_zero)
        rec = _power_records(plan, brightness_temp)[0]
        return [rec.copy(), plan.kcentres[0].copy()]

    def _own_plan(self, shape, bounds, lengths):
        return self._plan(shape, bounds, lengths, self.n_psbins, self.logk, self.ignore_kperp_zero,
                          self.ignore_kpar_zero, self.ignore_k_zero, self.device, self.edge_rule, self.kernel_path)

    def _geometry(self, ctx):
        """(boxes as one stack, plan) for a context.  Coeval: every redshift cube is a 1-chunk box."""
        cubes = ctx.get("brightness_temp")
        shape = tuple(cubes[0].shape)
        L = self.user_params.BOX_LEN
        plan = self._own_plan(shape, [(0, shape[2])], [(L, L, L)])
        return cubes, plan

    def _models_from_power(self, plan, rec):
        """rec: (nrec, out_per_box) power -> list over records of {"k","delta"} per chunk (likelihood.py:863)."""
        out = []
        for r in rec:
            m = []
            for c in range(plan.nchunks):
                k = plan.kcentres[c].copy()
                power = plan._chunk_view(r, c)
                m.append({"k": k, "delta": power * k**3 / (2 * np.pi**2)})
            out.append(m)
        return out

    def reduce_data(self, ctx):
        """One ``{"k", "delta"}`` dict per redshift (likelihood.py:639-654), all cubes in one device call."""
        cubes, plan = self._geometry(ctx)
        rec = _power_records(plan, list(cubes))
        return [m[0] for m in self._models_from_power(plan, rec)]

    def _tables(self, ks):
        """Data side of computeLikelihood at the given k arrays: (data, noise, mask) lists (likelihood.py:588-597)."""
        data, noise, mask = [], [], []
        for i, k in enumerate(ks):
            sel = np.logical_and(k <= self.max_k, k >= self.min_k)
            data.append(self.data_spline[i](k))
            noise.append(self.noise_spline[i](k) if self.noise_spline else np.zeros_like(k))
            mask.append(sel)
        return data, noise, mask

    def computeLikelihood(self, model):
        """ln L of one reduced model (likelihood.py:584-607); the chi^2 runs on the device."""
        ks = [np.asarray(m["k"], dtype=np.float64) for m in model]
        data, noise, mask = self._tables(ks)
        seg = np.array([len(k) for k in ks], dtype=np.int64)
        delta = np.concatenate([np.asarray(m["delta"], dtype=np.float64) for m in model])
        lnl = chi2_device(delta[None, :], np.concatenate(data), np.concatenate(noise), np.concatenate(mask), seg,
                          self.model_uncertainty, self.error_on_model, _device(self.device))[0]
        logger.debug(f"Likelihood computed: {lnl}")
        return float(lnl)

    def store(self, model, storage):
        """Blobs ``k_z{z}`` / ``delta_z{z}`` (likelihood.py:656-664)."""
        for i, m in enumerate(model):
            storage.update({k + "_z%s" % self.redshift[i]: v for k, v in m.items()})

    # ------------------------------------------------------------------ batched walkers
    def __getstate__(self):
        st = self.__dict__.copy()
        st.pop("_plan_tables", None)   # per-process cache (holds plan handles): not pickled into pool workers
        return st

    def _prepare_plan_data(self, plan):
        """Upload this likelihood's data / noise / k-mask tables at the plan's bin centres (walker independent).

        The spline evaluations are cached per plan, and the plan skips the upload when the tables already on the
        device are the same (``PowerPlan.set_data``); callers hold ``plan._lock`` across this and the evaluation."""
        cache = self.__dict__.setdefault("_plan_tables", {})
        hit = cache.get(id(plan))
        if hit is None or hit[0] is not plan:
            ks = [plan.kcentres[0]] * self._n_expected() if self._variant == "coeval" else list(plan.kcentres)
            data, noise, mask = self._tables(ks)
            nsets = self._n_expected() if self._variant == "coeval" else 1
            shape = (nsets, plan.nchunks, plan.nbins)
            hit = (plan, np.array(data).reshape(shape), np.array(noise).reshape(shape),
                   np.array(mask, dtype=np.uint8).reshape(shape))
            cache.clear()   # one geometry at a time is the rule; do not pin old plans
            cache[id(plan)] = hit
        plan.set_data(hit[1], hit[2], hit[3], self.model_uncertainty, self.error_on_model)

    def reduce_data_batch(self, ctxs):
        """``[reduce_data(ctx) for ctx in ctxs]`` with a single device call."""
        return self.likelihood_batch(ctxs, want_lnl=False)[1]

    def likelihood_batch(self, ctxs, want_lnl=True):
        """(lnL array, list of models) for many contexts: one batched FFT + binning + chi^2 call.

        Host lightcones / cubes are handed to the library by pointer (``ps21_lnl_host_ptrs``): nothing is stacked
        or concatenated on the host, and page-locked arrays (``plan.pinned_empty``) copy at the PCIe line rate."""
        if not ctxs:
            return np.zeros(0), []
        boxes, plan, per = [], None, 1
        for ctx in ctxs:
            cubes, pl = self._geometry(ctx)
            if plan is not None and pl is not plan:
                raise ValueError("all contexts of a batch must share one box geometry")
            plan, per = pl, len(cubes)
            boxes.extend(cubes)
        on_device = _is_cuda_tensor(boxes[0])
        if want_lnl:
            with plan._lock:   # data tables and evaluation belong together (plans are shared between likelihoods)
                self._prepare_plan_data(plan)
                if on_device:
                    import torch

                    t = torch.stack([b.contiguous() for b in boxes])
                    delta = torch.empty((t.shape[0], plan.out_per_box), dtype=torch.float64, device=t.device)
                    lnl = plan.lnl_device(t, delta=delta).cpu().numpy()
                    delta = delta.cpu().numpy()
                else:
                    lnl, delta, _ = plan.lnl_host(boxes, want_delta=True)
            lnl = lnl.reshape(len(ctxs), per).sum(axis=1)
            recs = [[{"k": plan.kcentres[c].copy(), "delta": plan._chunk_view(r, c).copy()}
                     for c in range(plan.nchunks)] for r in delta]
        else:
            lnl = None
            recs = self._models_from_power(plan, _power_records(plan, boxes))
        models = []
        for w in range(len(ctxs)):
            rows = recs[w * per:(w + 1) * per]
            models.append([r[0] for r in rows] if self._variant == "coeval" else rows[0])
        return lnl, models


class Likelihood1DPowerLightcone(Likelihood1DPowerCoeval):
    """The same likelihood on a lightcone cut into ``nchunks`` line-of-sight chunks (likelihood.py:685-892)."""

    required_cores = (core.LIGHTCONE_CORES,)
    _variant = "lightcone"

    def __init__(self, *args, datafile="", nchunks=1, **kwargs):
        super().__init__(*args, **kwargs)
        self.nchunks = nchunks
        self.datafile = [datafile] if isinstance(datafile, (str, Path)) else datafile

    @classmethod
    def from_builtin_data(cls, datafile="", **kwargs):
        """Create the class loading in built-in data (likelihood.py:702-707): ``<data dir>/<datafile>.npz``."""
        return cls(datafile=path.join(builtin_data_dir(), datafile + ".npz"), **kwargs)

    def _n_expected(self):
        return self.nchunks

    def setup(self):
        LikelihoodBaseFile.setup(self)
        if not self._use_data:
            return
        if len(self.data) != self.nchunks:
            raise ValueError("There needs to be one dataset (datafile) for each chunk!!")
        if self.noise and len(self.noise) != self.nchunks:
            raise ValueError("There needs to be one dataset (noisefile) for each chunk!")
        self._check_data_format()
        self._check_edge_rule()
        if self.noise:
            self._check_noise_format()

    @cached_property
    def paired_core(self):
        paired = [c for c in self._cores if isinstance(c, core.LIGHTCONE_CORES) and c.name == self.name]
        if len(paired) > 1:
            raise ValueError("You've got more than one CoreLightConeModule with the same name -- "
                             "they will overwrite each other!")
        return paired[0]

    @staticmethod
    def compute_power(box, length, n_psbins, log_bins=True, ignore_kperp_zero=True, ignore_kpar_zero=False,
                      ignore_k_zero=False):
        """Spherically averaged P(k) of one lightcone chunk: ``[power, k_centres]`` (likelihood.py:734-800).

        ``box`` may be a non-contiguous view ``bt[:, :, start:end]`` (likelihood.py:855); it is staged as a
        contiguous (n0, n1, nz) box.
        """
        shape = tuple(box.shape)
        plan = Likelihood1DPowerCoeval._plan(shape, [(0, shape[2])], [tuple(length)], n_psbins, log_bins,
                                             ignore_kperp_zero, ignore_kpar_zero, ignore_k_zero)
        if _is_cuda_tensor(box):
            box = box.contiguous()
        rec = _power_records(plan, box)[0]
        return [rec.copy(), plan.kcentres[0].copy()]

    @staticmethod
    def compute_power_2d(box, length, n_psbins, log_bins=False, ignore_kperp_zero=True, ignore_kpar_zero=False,
                         ignore_k_zero=False):
        """Cylindrical P(k_perp, k_par) = ``get_power(..., res_ndim=2)``.

        Returns ``[P2D (n_psbins, nz) in shifted k_par order, k_perp centres, k_par (nz,) shifted]``.
        """
        shape = tuple(box.shape)
        plan = Likelihood1DPowerCoeval._plan(shape, [(0, shape[2])], [tuple(length)], n_psbins, log_bins,
                                             ignore_kperp_zero, ignore_kpar_zero, ignore_k_zero,
                                             mode="cylindrical")
        if _is_cuda_tensor(box):
            box = box.contiguous()
        half = plan._chunk_view(_power_records(plan, box)[0], 0)  # (nb, nh)
        nz = shape[2]
        f = np.fft.fftshift(np.fft.fftfreq(nz, d=1.0 / nz)).round().astype(int)  # shifted integer frequencies
        cols = np.minimum(np.abs(f), nz - np.abs(f))
        kpar = np.fft.fftshift(np.fft.fftfreq(nz, d=float(length[2]) / nz)) * 2 * np.pi
        return [half[:, cols], plan.kcentres[0].copy(), kpar]

    def _geometry(self, ctx):
        lc = ctx.get("lightcone")
        bt = lc.brightness_temp
        shape = tuple(bt.shape)
        ci = binning.chunk_indices(lc.n_slices, self.nchunks)
        # IndexError for too few chunk starts, like the reference loop (likelihood.py:849-851)
        bounds = [(ci[i], ci[i + 1]) for i in range(self.nchunks)]
        L = self.user_params.BOX_LEN
        lengths = [(L, L, (b - a) * lc.cell_size) for a, b in bounds]
        return [bt], self._own_plan(shape, bounds, lengths)

    def reduce_data(self, ctx):
        """One ``{"k", "delta"}`` dict per chunk (likelihood.py:833-865), all chunks in one device call."""
        boxes, plan = self._geometry(ctx)
        rec = _power_records(plan, list(boxes))
        return self._models_from_power(plan, rec)[0]

    def store(self, model, storage):
        """Blobs ``k_{i}`` / ``delta_{i}`` (likelihood.py:867-876)."""
        for i, m in enumerate(model):
            storage.update({k + "_%s" % i: v for k, v in m.items()})


class Likelihood1DPowerLightconeUpper(Likelihood1DPowerLightcone):
    """Upper-limit (HERA-style) likelihood on the GPU spectra of a lightcone.

    The arithmetic is that of the reference's ``Likelihood1DPowerLightconeUpper.computeLikelihood``
    (likelihood.py:2249-2335): window-function matrix times the model Delta^2, 20 % model error, one-sided
    Gaussian (erf) likelihood per (band, field, k).  The reference feeds it from the emulator; here the model
    comes from the lightcone chunks: chunk ``band_chunks[b]`` is the spectrum of band ``z_bands[b]`` and is
    interpolated (linear in k, like ``data_spline``) onto that band's window-function k values.

    ``datafile``: ``.npz`` in the reference's layout -- ``z_bands``, ``band{z}`` (nfields, nk, 3) = [k, limit,
    variance] with NaN-padded k, ``wfband{z}`` (nfields, nk, nkwf), ``kwfband{z}`` (nkwf,).
    """

    _entry = "ps21_upper_limit_host"

    def __init__(self, *args, band_chunks=None, model_error_fraction=0.2, **kwargs):
        kwargs.setdefault("simulate", False)
        super().__init__(*args, **kwargs)
        self.band_chunks = band_chunks
        self.model_error_fraction = model_error_fraction

    def setup(self):
        LikelihoodBase.setup(self)
        self.data = self._read_data()
        self.redshifts = np.asarray(self.data[0]["z_bands"])
        self.k = [np.asarray(self.data[0]["kwfband" + str(int(np.round(z)))]) for z in self.redshifts]
        self.k_len = max(len(i) for i in self.k)
        if self.band_chunks is None:
            self.band_chunks = list(range(len(self.redshifts)))
        if len(self.band_chunks) != len(self.redshifts) or max(self.band_chunks) >= self.nchunks:
            raise ValueError("band_chunks must name one lightcone chunk per band")

    def _to_bands(self, chunk_models):
        """Interpolate per-chunk {"k","delta"} onto the window-function k of every band: (nbands, k_len)."""
        out = np.zeros((len(self.redshifts), self.k_len))
        for b, c in enumerate(self.band_chunks):
            m = chunk_models[c]
            ok = np.isfinite(m["delta"])
            spl = InterpolatedUnivariateSpline(m["k"][ok], m["delta"][ok], k=1)
            out[b, : len(self.k[b])] = spl(self.k[b])
        return out

    def reduce_data(self, ctx):
        chunk_models = super().reduce_data(ctx)
        return [{"k": self.k, "delta": self._to_bands(chunk_models)[None], "chunks": chunk_models}]

    def reduce_data_batch(self, ctxs):
        """``[reduce_data(ctx) for ctx in ctxs]``: the chunk spectra of all walkers come from one device call."""
        return self.likelihood_batch(ctxs, want_lnl=False)[1]

    def likelihood_batch(self, ctxs, want_lnl=True):
        """(lnL array, per-walker models) -- the batched route of ``chain.call_batch`` / ``BatchedWalkerPool``.

        The Gaussian chi^2 of the parent class does not apply here (the data file holds limits and window
        functions, no ``k`` / ``delta``): chunk spectra come from the parent's batched reduction, the likelihood
        from this class's own ``computeLikelihood`` on all walkers at once."""
        if not ctxs:
            return np.zeros(0), []
        chunk_models = Likelihood1DPowerLightcone.likelihood_batch(self, ctxs, want_lnl=False)[1]
        delta = np.stack([self._to_bands(m) for m in chunk_models])          # (N, nbands, k_len)
        models = [[{"k": self.k, "delta": delta[i][None], "chunks": chunk_models[i]}] for i in range(len(ctxs))]
        lnl = None
        if want_lnl:
            lnl = np.atleast_1d(np.asarray(self.computeLikelihood([{"k": self.k, "delta": delta}]), dtype=float))
        return lnl, models

    def computeLikelihood(self, model):
        """lnL of every model in ``model[0]["delta"]`` (N, nbands, k_len); the matvec + erf run on the device."""
        import ctypes as C

        lib = _capi.load()
        delta = np.asarray(model[0]["delta"], dtype=np.float64)
        N = delta.shape[0]
        lnl = np.zeros(N)
        dp = C.POINTER(C.c_double)
        hd = self.data[0]
        err_all = model[0].get("delta_err")
        for band in self.redshifts:
            key = str(round(float(band)))
            zbin = int(np.argmin(abs(band - self.redshifts)))
            nkwf = len(hd["kwfband" + key])
            m = np.ascontiguousarray(delta[:, zbin, :nkwf])
            err = None if err_all is None else np.ascontiguousarray(np.asarray(err_all)[zbin, :nkwf], dtype=np.float64)
            for field in range(hd["band" + key].shape[0]):
                ks = hd["band" + key][field, :, 0]
                nk = int(np.sum(~np.isnan(ks)))
                vals = np.ascontiguousarray(hd["band" + key][field, :nk, 1], dtype=np.float64)
                var = np.ascontiguousarray(hd["band" + key][field, :nk, 2], dtype=np.float64)
                wf = np.ascontiguousarray(hd["wfband" + key][field, :nk, :].reshape(nk, nkwf), dtype=np.float64)
                _capi.check(getattr(lib, self._entry)(
                    _device(self.device), N, nk, nkwf, m.ctypes.data_as(dp), wf.ctypes.data_as(dp),
                    vals.ctypes.data_as(dp), var.ctypes.data_as(dp),
                    err.ctypes.data_as(dp) if err is not None else None, float(self.model_error_fraction),
                    lnl.ctypes.data_as(dp)))
        return lnl if N > 1 else float(lnl[0])

    def store(self, model, storage):
        super().store(model[0]["chunks"] if isinstance(model[0]["chunks"][0], dict) else model[0]["chunks"][0], storage)


class Likelihood1DPowerLightconeWindowed(Likelihood1DPowerLightconeUpper):
    """Measured (not upper-limit) band powers seen through window functions: the Gaussian branch of the
    reference's ``Likelihood1DPowerCoeval.computeLikelihood`` (likelihood.py:517-583) -- same data layout, window
    matvec, 20 % model error and optional emulator-style ``delta_err`` term as the upper-limit class, with
    ``-0.5 (W.model - value)^2 / err^2`` instead of the one-sided erf."""

    _entry = "ps21_windowed_gauss_host"


def chi2_device(delta, data, noise, mask, seg, model_uncertainty, error_on_model, device=0):
    """-0.5 * sum_seg sum_k (delta-data)^2 / ((mu*delta|data)^2 + noise^2) for every row of ``delta``."""
    import ctypes as C

    lib = _capi.load()
    delta = np.ascontiguousarray(delta, dtype=np.float64)
    nrec, n = delta.shape
    data = np.ascontiguousarray(data, dtype=np.float64)
    noise = np.ascontiguousarray(noise, dtype=np.float64)
    mask = np.ascontiguousarray(mask, dtype=np.uint8)
    seg = np.ascontiguousarray(seg, dtype=np.int64)
    out = np.empty(nrec, dtype=np.float64)
    dp = C.POINTER(C.c_double)
    _capi.check(lib.ps21_chi2_host(int(device), nrec, n, delta.ctypes.data_as(dp), data.ctypes.data_as(dp),
                                   noise.ctypes.data_as(dp), mask.ctypes.data_as(C.POINTER(C.c_uint8)),
                                   len(seg), seg.ctypes.data_as(C.POINTER(C.c_int64)),
                                   float(model_uncertainty), int(bool(error_on_model)),
                                   out.ctypes.data_as(dp)))
    return out
