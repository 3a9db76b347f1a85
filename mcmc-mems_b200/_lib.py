"""ctypes binding of ``libmems_b200.so`` (C ABI: include/mems_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a).  There is
no CPU fallback: if the library is missing or no sm_100 device is present every compute entry
point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MEMS_LIB") or os.path.join(_HERE, "libmems_b200.so")   # MEMS_LIB: experimental build

MEMS_PMAX = 4
MEMS_WIDTH = 64
MEMS_HIDDEN = 6
PATH_FP32 = 0
PATH_TC = 1
SAMPLER_MH = 0
SAMPLER_CWMH = 1
SAMPLER_DA = 2

EXPORTS = [
    "mems_last_error", "mems_version", "mems_create", "mems_destroy", "mems_set_weights",
    "mems_set_problem", "mems_forward", "mems_init_chains", "mems_run", "mems_run_explicit",
    "mems_draw", "mems_get_state", "mems_get_aux", "mems_set_sampler", "mems_mlp_forward", "mems_chain_moments", "mems_launch_count", "mems_rows_evaluated", "mems_lsq", "mems_chain_autocov", "mems_set_coarse_fft",
    "mems_sort_capacity", "mems_sort_f64", "mems_rank_normalize", "mems_ess_geyer",
]


class MemsConfig(C.Structure):
    _fields_ = [("n_params", C.c_int32), ("n_time", C.c_int32), ("n_chains", C.c_int32),
                ("path", C.c_int32), ("device", C.c_int32), ("chains_per_block", C.c_int32),
                ("reserved", C.c_int32 * 2)]


class MemsError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the shared library once; raise loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MemsError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    L.mems_last_error.restype = C.c_char_p
    L.mems_last_error.argtypes = []
    L.mems_version.restype = C.c_int
    L.mems_create.argtypes = [C.POINTER(MemsConfig), C.POINTER(vp)]
    L.mems_destroy.argtypes = [vp]
    L.mems_destroy.restype = None
    L.mems_set_weights.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.mems_set_problem.argtypes = [vp, vp, vp, vp, vp, dbl, vp, vp, vp]
    L.mems_forward.argtypes = [vp, vp, i32, vp, vp, vp]
    L.mems_init_chains.argtypes = [vp, vp, dbl, u64, u64, vp]
    L.mems_run.argtypes = [vp, i32, i32, i32, vp, vp, vp]
    L.mems_run_explicit.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp]
    L.mems_draw.argtypes = [vp, u64, i32, vp, vp, vp]
    L.mems_get_state.argtypes = [vp, vp, vp, vp, vp, C.POINTER(u64), vp]
    L.mems_get_aux.argtypes = [vp, vp, vp, vp, vp]
    L.mems_set_sampler.argtypes = [vp, i32, vp, i32]
    L.mems_mlp_forward.argtypes = [i32, i32, C.POINTER(i32), C.POINTER(vp), C.POINTER(vp), vp, i64, vp, vp]
    L.mems_chain_moments.argtypes = [i32, vp, i32, i32, i32, vp, vp]
    L.mems_launch_count.argtypes = [vp]
    L.mems_launch_count.restype = i64
    L.mems_rows_evaluated.argtypes = [vp, C.POINTER(u64), vp]
    L.mems_lsq.argtypes = [vp, vp, i32, vp, i32, vp, vp, i32, dbl, dbl, dbl, vp, vp, vp, vp, vp]
    L.mems_chain_autocov.argtypes = [i32, vp, i32, i32, i32, i32, vp, vp]
    L.mems_set_coarse_fft.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), vp, vp]
    L.mems_sort_capacity.argtypes = [i64]
    L.mems_sort_capacity.restype = i64
    L.mems_sort_f64.argtypes = [i32, vp, i64, vp]
    L.mems_rank_normalize.argtypes = [i32, vp, i64, vp, i64, vp, vp]
    L.mems_ess_geyer.argtypes = [i32, vp, i32, i32, i64, i64, vp, vp, vp]
    _lib = L
    return L


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().mems_last_error().decode("utf-8", "replace")
        raise MemsError(f"{what} failed (code {rc}): {msg}")


# -- development tools (tools/libmems_b200_tools.so; not part of the product ABI) -----------------
TOOLS_LIB_PATH = os.environ.get("MEMS_TOOLS_LIB") or os.path.join(os.path.dirname(_HERE), "tools", "libmems_b200_tools.so")
_tools = None


def tools_lib():
    """Micro-benchmarks and the UMMA self-test (tools/mems_b200_tools.h)."""
    global _tools
    if _tools is not None:
        return _tools
    if not os.path.exists(TOOLS_LIB_PATH):
        raise MemsError(f"{TOOLS_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(TOOLS_LIB_PATH)
    vp, i32 = C.c_void_p, C.c_int32
    L.mems_tools_last_error.restype = C.c_char_p
    L.mems_tools_selftest_umma.argtypes = [i32, i32, vp, vp, vp, vp]
    L.mems_tools_microbench.argtypes = [i32, i32, i32, i32, i32, vp, vp, vp]
    _tools = L
    return L


def check_tools(rc: int, what: str):
    if rc != 0:
        msg = tools_lib().mems_tools_last_error().decode("utf-8", "replace")
        raise MemsError(f"{what} failed (code {rc}): {msg}")
