"""B200-native power-spectrum likelihood for 21CMMC (drop-in for the PS-likelihood hot path).

Import as ``py21cmmc_b200`` (see ../py21cmmc_b200/__init__.py).  Layout:

* ``csrc/``        hand-written sm_100a CUDA kernels + the C ABI of ``include/ps21.h`` (libps21.so)
* ``_capi``        ctypes binding, ``plan`` per-process plan cache, ``binning`` host geometry
* ``likelihood``   ``Likelihood1DPowerCoeval`` / ``Likelihood1DPowerLightcone`` with the reference's API
* ``core`` / ``chain``  the plugin base classes and a minimal chain/context the likelihoods attach to
"""
from . import binning
from ._capi import Ps21Error

__version__ = "0.1.0"
__all__ = ["binning", "Ps21Error"]
