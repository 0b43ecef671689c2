"""ctypes binding of libbagpu.so -- the same symbols the Julia glue binds with ``ccall``
(include/bagpu.h).  There is no CPU fallback: if the CUDA library is missing or a
call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("BAGPU_LIB", os.path.join(HERE, "libbagpu.so"))  # BAGPU_LIB: A/B builds

BA_OK, BA_ERR_ARG, BA_ERR_CUDA, BA_ERR_UNSORTED, BA_ERR_COMM, BA_ERR_NUMERIC = range(6)
SOLVERS = {"auto": 0, "pcg": 1, "exact": 2, "mixed": 3}  # BA_SOLVER_*
SOLVER_NAMES = {v: k for k, v in SOLVERS.items()}
_CODE = {1: "bad argument", 2: "CUDA failure", 3: "observations not point-major", 4: "NCCL failure",
         5: "numeric breakdown"}


class BAError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libbagpu: %s (code %d): %s" % (_CODE.get(code, "error"), code, msg))
        self.code = code


class LMParams(C.Structure):
    """ba_lm_params: keyword arguments of Levenberg_Marquardt (src/lm.jl:15-26) + PCG controls."""
    _fields_ = [(n, C.c_double) for n in ("restol", "satol", "srtol", "oatol", "ortol", "atol", "rtol",
                                          "nu_d", "nu_m", "lam", "delta_d")] + \
               [("ite_max", C.c_int64), ("linesearch", C.c_int32), ("pcg_max_iter", C.c_int32),
                ("pcg_tol", C.c_double), ("solver", C.c_int32), ("reserved", C.c_int32)]


class LMRow(C.Structure):
    """ba_lm_row: one log_row of src/lm.jl:304."""
    _fields_ = [("iter", C.c_int64)] + [(n, C.c_double) for n in ("f", "df", "dfeas", "lam", "delta_norm", "rho")] + \
               [("accepted", C.c_int32), ("acc_str", C.c_int32), ("pcg_iters", C.c_int32), ("ntimes", C.c_int32),
                ("solver", C.c_int32), ("converged", C.c_int32), ("solve_rel", C.c_double)]


class LMStats(C.Structure):
    _fields_ = [("status", C.c_int32), ("pad", C.c_int32), ("iter", C.c_int64), ("objective", C.c_double),
                ("dual_feas", C.c_double), ("lambda_final", C.c_double), ("elapsed_s", C.c_double),
                ("pcg_iters_total", C.c_int64), ("t_eval_ms", C.c_double), ("t_assemble_ms", C.c_double),
                ("t_pcg_ms", C.c_double), ("t_backsub_ms", C.c_double), ("capped_solves", C.c_int64),
                ("worst_solve_rel", C.c_double), ("t_prepare_ms", C.c_double), ("t_schur_ms", C.c_double),
                ("t_chol_ms", C.c_double), ("chol_n", C.c_int64), ("chol_count", C.c_int64),
                ("mixed_fallbacks", C.c_int64)]


ITER_CB = C.CFUNCTYPE(None, C.POINTER(LMRow), C.c_void_p)

# every symbol include/bagpu.h declares: name -> (restype, argtypes)
_vp, _i64, _i32, _f64 = C.c_void_p, C.c_int64, C.c_int32, C.c_double
SYMBOLS = {
    "ba_create": (C.c_int, [_i64, _i64, _i64, _vp, _vp, _vp, C.c_int, C.POINTER(_vp)]),
    "ba_create_sharded": (C.c_int, [_i64, _i64, _i64, _vp, _vp, _vp, C.c_int, C.c_int, C.c_int, C.POINTER(_vp)]),
    "ba_create_multi": (C.c_int, [_i64, _i64, _i64, _vp, _vp, _vp, C.c_int, _vp, C.POINTER(_vp)]),
    "ba_destroy": (C.c_int, [_vp]),
    "ba_shard_range": (C.c_int, [_vp] + [C.POINTER(_i64)] * 4),
    "ba_partition_observations": (C.c_int, [_i64, _vp, C.c_int, _vp]),
    "ba_last_error": (C.c_char_p, [_vp]),
    "ba_version": (C.c_char_p, []),
    "ba_measure_fp64_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "ba_measure_fp64_mma_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double)]),
    "ba_set_stream": (C.c_int, [_vp, _vp]),
    "ba_alloc_pinned": (C.c_int, [C.c_uint64, C.POINTER(_vp)]),
    "ba_free_pinned": (C.c_int, [_vp]),
    "ba_residual": (C.c_int, [_vp, _vp, _vp]),
    "ba_jac_structure": (C.c_int, [_vp, _vp, _vp]),
    "ba_jac_coord": (C.c_int, [_vp, _vp, _vp]),
    "ba_residual_jac": (C.c_int, [_vp, _vp, _vp, _vp]),
    "ba_jprod": (C.c_int, [_vp, _vp, _vp, _vp]),
    "ba_jtprod": (C.c_int, [_vp, _vp, _vp, _vp]),
    "ba_residual_dev": (C.c_int, [_vp, _vp, _vp]),
    "ba_jac_structure_dev": (C.c_int, [_vp, _vp, _vp]),
    "ba_jac_coord_dev": (C.c_int, [_vp, _vp, _vp]),
    "ba_residual_jac_dev": (C.c_int, [_vp, _vp, _vp, _vp]),
    "ba_jprod_dev": (C.c_int, [_vp, _vp, _vp, _vp]),
    "ba_jtprod_dev": (C.c_int, [_vp, _vp, _vp, _vp]),
    "ba_sync": (C.c_int, [_vp]),
    "ba_set_coarse_clusters": (C.c_int, [_vp, C.c_int]),
    "ba_set_profiling": (C.c_int, [_vp, C.c_int]),
    "ba_set_solver": (C.c_int, [_vp, C.c_int]),
    "ba_last_solve_info": (C.c_int, [_vp, C.POINTER(_i32), C.POINTER(_i32), C.POINTER(_f64), C.POINTER(_i32)]),
    "ba_dbg_chol": (C.c_int, [C.c_int, _i64, _vp, _vp, _vp, _vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "ba_dbg_probe_peak": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "ba_dbg_chol32": (C.c_int, [C.c_int, _i64, _vp, _vp, _vp, _vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]),
    "ba_last_eval_ms": (C.c_int, [_vp, C.POINTER(C.c_float)]),
    "ba_lm_default_params": (None, [C.POINTER(LMParams)]),
    "ba_lm_step": (C.c_int, [_vp, _vp, _f64, _f64, _i32, _vp, C.POINTER(_f64), C.POINTER(_f64), _vp,
                             C.POINTER(_i32)]),
    "ba_lm_solve": (C.c_int, [_vp, _vp, C.POINTER(LMParams), C.POINTER(LMStats), _vp, _vp]),
    "ba_set_deflation": (C.c_int, [_vp, C.c_int]),
    "ba_dbg_tridiag_eig": (C.c_int, [_vp, _vp, _i32, _vp, _vp]),
    "ba_dbg_tridiag_smallest": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp, _vp]),
    "ba_dbg_select_columns": (C.c_int, [_vp, _i32, _i32, _f64, _vp, C.POINTER(_i32)]),
    "ba_comm_unique_id": (C.c_int, [_vp]),
    "ba_comm_init": (C.c_int, [_vp, _vp]),
    "ba_comm_ipc_export": (C.c_int, [_vp, _vp]),
    "ba_comm_ipc_import": (C.c_int, [_vp, _vp]),
    "ba_comm_ipc_disable": (C.c_int, [_vp]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libbagpu.so.  Raises if it has not been built (python -m ... build / __graft_entry__.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("libbagpu.so is not built: run `python __graft_entry__.py build` "
                              "(there is no CPU fallback for the CUDA path)")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)  # AttributeError if the library lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int, handle=None):
    if rc != BA_OK:
        msg = lib().ba_last_error(handle) if handle else b""
        raise BAError(rc, (msg or b"").decode("utf-8", "replace"))
