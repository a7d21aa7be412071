"""ctypes binding of libps21.so (include/ps21.h).  No torch types cross this boundary.

The library is the hand-written sm_100a implementation; there is no CPU fallback: if the shared
object is missing or a CUDA call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# PS21_LIBRARY selects another build of the same library (A/B comparisons of kernel variants)
LIB_PATH = os.environ.get("PS21_LIBRARY") or os.path.join(HERE, "libps21.so")

PS21_MODE_SPHERICAL = 0
PS21_MODE_CYLINDRICAL = 1
PS21_PATH_AUTO = 0
PS21_PATH_GENERAL = 1
PS21_PATH_FUSED = 2
PS21_PATH_TWOPASS = 3

EXPORTS = (
    "ps21_last_error", "ps21_version", "ps21_plan_create", "ps21_plan_destroy", "ps21_plan_info",
    "ps21_plan_set_profiling", "ps21_plan_profile", "ps21_plan_layout", "ps21_plan_counts", "ps21_power_batched", "ps21_set_data", "ps21_lnl_batched",
    "ps21_power_host", "ps21_lnl_host", "ps21_lnl_from_delta", "ps21_chi2_host", "ps21_upper_limit_host",
    "ps21_windowed_gauss_host", "ps21_power_host_ptrs", "ps21_lnl_host_ptrs", "ps21_host_alloc", "ps21_host_free",
    "ps21_host_register", "ps21_host_unregister", "ps21_plan_set_graphs", "ps21_plan_graph_stats",
    "ps21_plan_release_workspace",
)


class Ps21Error(RuntimeError):
    pass


class Desc(C.Structure):
    _fields_ = [
        ("n0", C.c_int32), ("n1", C.c_int32), ("n2_total", C.c_int32), ("nchunks", C.c_int32),
        ("chunk_start", C.POINTER(C.c_int32)), ("chunk_end", C.POINTER(C.c_int32)),
        ("volume", C.POINTER(C.c_double)),
        ("k0sq", C.POINTER(C.c_double)), ("k1sq", C.POINTER(C.c_double)), ("k2sq", C.POINTER(C.c_double)),
        ("nbins", C.c_int32),
        ("edges", C.POINTER(C.c_double)), ("kcentres", C.POINTER(C.c_double)),
        ("ignore_kperp_zero", C.c_int32), ("ignore_kpar_zero", C.c_int32), ("ignore_k_zero", C.c_int32),
        ("mode", C.c_int32), ("device", C.c_int32), ("path", C.c_int32),
        ("workspace_bytes", C.c_int64),
    ]


class Info(C.Structure):
    _fields_ = [
        ("nshapes", C.c_int32), ("path_of_chunk0", C.c_int32), ("out_per_box", C.c_int64),
        ("workspace_bytes", C.c_int64), ("alg_bytes_per_box", C.c_int64), ("launches_last_call", C.c_int64),
        ("dominant_ms_last", C.c_double),
    ]


_lib = None
_lib_path = None


def _declare(lib):
    vp, i64, dp = C.c_void_p, C.c_int64, C.POINTER(C.c_double)
    lib.ps21_last_error.restype = C.c_char_p
    lib.ps21_last_error.argtypes = []
    lib.ps21_version.restype = C.c_int
    lib.ps21_plan_create.argtypes = [C.POINTER(Desc), C.POINTER(vp)]
    lib.ps21_plan_destroy.argtypes = [vp]
    lib.ps21_plan_info.argtypes = [vp, C.POINTER(Info)]
    lib.ps21_plan_set_profiling.argtypes = [vp, C.c_int32]
    lib.ps21_plan_profile.argtypes = [vp, dp, C.POINTER(i64)]
    lib.ps21_plan_layout.argtypes = [vp, C.POINTER(i64)]
    lib.ps21_plan_counts.argtypes = [vp, C.POINTER(i64)]
    lib.ps21_power_batched.argtypes = [vp, vp, i64, i64, vp, vp]
    lib.ps21_set_data.argtypes = [vp, C.c_int32, dp, dp, C.POINTER(C.c_uint8), C.c_double, C.c_int32]
    lib.ps21_lnl_batched.argtypes = [vp, vp, i64, i64, vp, vp, vp, vp]
    lib.ps21_power_host.argtypes = [vp, vp, i64, i64, vp]
    lib.ps21_lnl_host.argtypes = [vp, vp, i64, i64, vp, vp, vp]
    lib.ps21_lnl_from_delta.argtypes = [vp, dp, i64, dp]
    lib.ps21_chi2_host.argtypes = [C.c_int32, i64, i64, dp, dp, dp, C.POINTER(C.c_uint8), i64, C.POINTER(i64),
                                   C.c_double, C.c_int32, dp]
    lib.ps21_upper_limit_host.argtypes = [C.c_int32, i64, i64, i64, dp, dp, dp, dp, dp, C.c_double, dp]
    lib.ps21_windowed_gauss_host.argtypes = lib.ps21_upper_limit_host.argtypes
    lib.ps21_power_host_ptrs.argtypes = [vp, C.POINTER(vp), i64, vp]
    lib.ps21_lnl_host_ptrs.argtypes = [vp, C.POINTER(vp), i64, vp, vp, vp]
    lib.ps21_host_alloc.argtypes = [C.POINTER(vp), i64]
    lib.ps21_host_free.argtypes = [vp]
    lib.ps21_host_register.argtypes = [vp, i64]
    lib.ps21_host_unregister.argtypes = [vp]
    lib.ps21_plan_set_graphs.argtypes = [vp, C.c_int32]
    lib.ps21_plan_graph_stats.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    lib.ps21_plan_release_workspace.argtypes = [vp]
    for name in EXPORTS:
        if name not in ("ps21_last_error",):
            getattr(lib, name).restype = C.c_int


def _testing_set_library(path: str | None):
    """Test hook: make ``load()`` resolve to another build of the same sources (tests/emu)."""
    global _lib, _lib_path, LIB_PATH
    _lib = None
    _lib_path = None
    LIB_PATH = path or os.path.join(HERE, "libps21.so")


def load(path: str | None = None):
    """Load (once) and return the shared library.  Raises Ps21Error when it is not built."""
    global _lib, _lib_path
    path = path or LIB_PATH
    if _lib is not None and _lib_path == path:
        return _lib
    if not os.path.exists(path):
        raise Ps21Error(
            f"{path} not found: build it with `python 21cmmc_b200/build.py` "
            "(this package has no CPU fallback)")
    lib = C.CDLL(path)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise Ps21Error(f"{path} does not export {name}")
    _declare(lib)
    _lib, _lib_path = lib, path
    return lib


def check(rc: int):
    if rc != 0:
        msg = load(_lib_path).ps21_last_error()
        raise Ps21Error(f"libps21 error {rc}: {msg.decode() if msg else ''}")
