"""Importable alias of the package that lives in ``21cmmc_b200/``.

The product directory is named after the reference (``21cmmc_b200``), which is not a valid Python
identifier; ``import py21cmmc_b200`` loads that directory's ``__init__.py`` under this name so
that sub-modules resolve as ``py21cmmc_b200.likelihood`` etc.  No code lives here.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_root = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "21cmmc_b200")
_spec = _ilu.spec_from_file_location(__name__, _os.path.join(_root, "__init__.py"),
                                     submodule_search_locations=[_root])
_mod = _ilu.module_from_spec(_spec)
_sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
