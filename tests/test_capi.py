"""The C-ABI library loads and exports every symbol include/ps21.h declares (no compute without a GPU)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "ps21.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ps21_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    import py21cmmc_b200  # noqa: F401
    from py21cmmc_b200 import _capi

    import importlib.util

    spec = importlib.util.spec_from_file_location("ps21_build", os.path.join(ROOT, "21cmmc_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    lib_path = b.build()
    lib = ctypes.CDLL(lib_path)
    names = _declared()
    assert len(names) >= 13
    for n in names:
        assert hasattr(lib, n), f"{n} declared in ps21.h but not exported"
    assert set(_capi.EXPORTS) == set(names)
    lib.ps21_version.restype = ctypes.c_int
    assert lib.ps21_version() >= 100


def test_sass_is_sm100a():
    import subprocess

    lib = os.path.join(ROOT, "21cmmc_b200", "libps21.so")
    out = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_missing_library_fails_loudly(tmp_path):
    import py21cmmc_b200  # noqa: F401
    from py21cmmc_b200 import Ps21Error, _capi
    import pytest

    _capi._testing_set_library(str(tmp_path / "nope.so"))
    try:
        with pytest.raises(Ps21Error):
            _capi.load()
    finally:
        _capi._testing_set_library(None)
