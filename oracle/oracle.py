"""ctypes loader for the CPU oracle (oracle/ba_oracle.c).

TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs import this module; the product package never does.

All index arrays are 1-based int64 exactly as the reference's ``BALNLPModel`` holds them
(src/BALNLPModels.jl:79-88).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libba_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile oracle/ba_oracle.c -> oracle/libba_oracle.so (gcc, -ffp-contract=off)."""
    src = os.path.join(_HERE, "ba_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libba_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


class LMParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("restol", "satol", "srtol", "oatol", "ortol", "atol", "rtol",
                                          "nu_d", "nu_m", "lam", "delta_d")] + \
               [("ite_max", C.c_int64), ("linesearch", C.c_int32), ("nthreads", C.c_int32)]


class LMRow(C.Structure):
    _fields_ = [("iter", C.c_int64)] + [(n, C.c_double) for n in ("f", "df", "dfeas", "lam", "delta_norm", "rho")] + \
               [("accepted", C.c_int32), ("acc_str", C.c_int32)]


class LMStats(C.Structure):
    _fields_ = [("status", C.c_int32), ("iter", C.c_int64), ("objective", C.c_double), ("dual_feas", C.c_double),
                ("lambda_final", C.c_double), ("nrows", C.c_int64), ("ldl_nnz", C.c_int64)]


STATUS = {0: "unknown", 1: "small_step", 2: "first_order", 3: "small_residual", 4: "acceptable",
          5: "neg_pred", 6: "exception", 7: "max_iter"}


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.bao_rodrigues_rotation.argtypes = [_f64p, _f64p, _f64p]
        L.bao_scaling_factor.argtypes = [_f64p, C.c_double, C.c_double]
        L.bao_scaling_factor.restype = C.c_double
        L.bao_projection_jump.argtypes = [_f64p, _f64p, _f64p, C.c_double, C.c_double, C.c_double, _f64p]
        L.bao_residuals.argtypes = [_i64p, _i64p, _f64p, _f64p, C.c_int64, C.c_int64, C.c_int]
        L.bao_cons.argtypes = [_i64p, _i64p, _f64p, _f64p, _f64p, C.c_int64, C.c_int64, C.c_int]
        L.bao_jac_structure.argtypes = [_i64p, _i64p, C.c_int64, C.c_int64, _i64p, _i64p]
        L.bao_jac_coord.argtypes = [_i64p, _i64p, _f64p, _f64p, C.c_int64, C.c_int64, C.c_int]
        L.bao_cons_jac.argtypes = [_i64p, _i64p, _f64p, _f64p, _f64p, _f64p, C.c_int64, C.c_int64, C.c_int]
        L.bao_mul_sparse.argtypes = [_i64p, _i64p, _f64p, _f64p, C.c_int64, _f64p, C.c_int64]
        L.bao_ldl_solve_csc.argtypes = [C.c_int64, _i64p, _i64p, _f64p, C.c_void_p, _f64p]
        L.bao_ldl_solve_csc.restype = C.c_int
        L.bao_lm_default_params.argtypes = [C.POINTER(LMParams)]
        L.bao_lm_step.argtypes = [_i64p, _i64p, _f64p, C.c_int64, C.c_int64, C.c_int64, _f64p, C.c_double,
                                  _f64p, C.POINTER(C.c_double), C.c_void_p]
        L.bao_lm_step.restype = C.c_int
        L.bao_lm_solve.argtypes = [_i64p, _i64p, _f64p, C.c_int64, C.c_int64, C.c_int64, _f64p,
                                   C.POINTER(LMParams), C.POINTER(LMStats), C.c_void_p, C.c_int64]
        L.bao_lm_solve.restype = C.c_int
        L.bao_lm_solve_ex.argtypes = L.bao_lm_solve.argtypes + [C.c_int]
        L.bao_lm_solve_ex.restype = C.c_int
        L.bao_lm_step_schur.argtypes = [_i64p, _i64p, _f64p, C.c_int64, C.c_int64, C.c_int64, _f64p, C.c_double,
                                        C.c_int, _f64p, C.POINTER(C.c_double), C.c_void_p]
        L.bao_lm_step_schur.restype = C.c_int
        L.bao_schur_system.argtypes = [_i64p, _i64p, _f64p, _f64p, C.c_int64, C.c_int64, C.c_int64, C.c_double,
                                       C.c_int, _f64p, _f64p, _f64p, _f64p]
        L.bao_schur_backsub.argtypes = [_i64p, _i64p, _f64p, C.c_int64, C.c_int64, C.c_int64, _f64p, _f64p, _f64p,
                                        C.c_int, _f64p]
        L.bao_chol_solve.argtypes = [C.c_int64, _f64p, _f64p, C.c_int]
        L.bao_chol_solve.restype = C.c_int
        L.bao_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def rodrigues_rotation(r, x):
    out = np.empty(3)
    lib().bao_rodrigues_rotation(_f64(r), _f64(x), out)
    return out


def scaling_factor(p2, k1, k2):
    return lib().bao_scaling_factor(_f64(p2), float(k1), float(k2))


def projection_jump(X, r, t, f, k1, k2):
    out = np.empty(2)
    lib().bao_projection_jump(_f64(X), _f64(r), _f64(t), float(f), float(k1), float(k2), out)
    return out


def residuals(cam_idx, pnt_idx, x, nobs, npnts, nthreads=1):
    """residuals! (src/BALNLPModels.jl:39-55): projections only, pt2d NOT subtracted."""
    r = np.empty(2 * nobs)
    lib().bao_residuals(_i64(cam_idx), _i64(pnt_idx), _f64(x), r, nobs, npnts, nthreads)
    return r


def cons(cam_idx, pnt_idx, pt2d, x, npnts, nthreads=1):
    """cons! (src/BALNLPModels.jl:115-122)."""
    nobs = len(cam_idx)
    cx = np.empty(2 * nobs)
    lib().bao_cons(_i64(cam_idx), _i64(pnt_idx), _f64(pt2d), _f64(x), cx, nobs, npnts, nthreads)
    return cx


def jac_structure(cam_idx, pnt_idx, npnts):
    nobs = len(cam_idx)
    rows = np.empty(24 * nobs, dtype=np.int64)
    cols = np.empty(24 * nobs, dtype=np.int64)
    lib().bao_jac_structure(_i64(cam_idx), _i64(pnt_idx), nobs, npnts, rows, cols)
    return rows, cols


def jac_coord(cam_idx, pnt_idx, x, npnts, nthreads=1):
    nobs = len(cam_idx)
    vals = np.empty(24 * nobs)
    lib().bao_jac_coord(_i64(cam_idx), _i64(pnt_idx), _f64(x), vals, nobs, npnts, nthreads)
    return vals


def cons_jac(cam_idx, pnt_idx, pt2d, x, npnts, nthreads=1, cx=None, vals=None):
    nobs = len(cam_idx)
    cx = np.empty(2 * nobs) if cx is None else cx
    vals = np.empty(24 * nobs) if vals is None else vals
    lib().bao_cons_jac(cam_idx, pnt_idx, pt2d, x, cx, vals, nobs, npnts, nthreads)
    return cx, vals


def mul_sparse(rows, cols, vals, x, l):
    xr = np.empty(l)
    lib().bao_mul_sparse(_i64(rows), _i64(cols), _f64(vals), _f64(x), len(vals), xr, l)
    return xr


def ldl_solve_csc(n, Ap, Ai, Ax, b, P=None):
    b = _f64(b).copy()
    Pp = None if P is None else _i64(P).ctypes.data_as(C.c_void_p)
    keep = None if P is None else _i64(P)
    if keep is not None:
        Pp = keep.ctypes.data_as(C.c_void_p)
    rc = lib().bao_ldl_solve_csc(n, _i64(Ap), _i64(Ai), _f64(Ax), Pp, b)
    if rc:
        raise ArithmeticError("SQDException: zero pivot (src/ldl_aux.jl:199)")
    return b


def default_params(**kw) -> LMParams:
    p = LMParams()
    lib().bao_lm_default_params(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def lm_step(cam_idx, pnt_idx, pt2d, ncams, npnts, x, lam, want_jtr=False):
    """One damped solve (J'J + lam I) d = -J'r through the augmented LDL (src/lm.jl:68-100,175-229)."""
    nobs = len(cam_idx)
    nvar = 9 * ncams + 3 * npnts
    delta = np.empty(nvar)
    dr2 = C.c_double()
    jtr = np.empty(nvar) if want_jtr else None
    rc = lib().bao_lm_step(_i64(cam_idx), _i64(pnt_idx), _f64(pt2d), ncams, npnts, nobs, _f64(x), float(lam),
                           delta, C.byref(dr2), None if jtr is None else jtr.ctypes.data_as(C.c_void_p))
    if rc:
        raise ArithmeticError("SQDException")
    return (delta, dr2.value, jtr) if want_jtr else (delta, dr2.value)


@dataclass
class LMResult:
    status: str
    iter: int
    objective: float
    dual_feas: float
    lambda_final: float
    solution: np.ndarray
    log: list
    ldl_nnz: int


def lm_step_schur(cam_idx, pnt_idx, pt2d, ncams, npnts, x, lam, want_jtr=False, nthreads=None, dense="auto"):
    """The same damped solve as lm_step, eliminated in the order a fill-reducing permutation (AMD/Metis,
    src/lm.jl:85-87) takes on a BA Jacobian: residual rows, 3x3 point blocks, then a dense Cholesky of the
    reduced camera system.  Finishes at every BASELINE.json size up to Venice-1778 (S is 16002^2 there).
    dense = "c" (oracle's own blocked Cholesky), "scipy" (LAPACK dpotrf, faster for big S) or "auto"."""
    nobs = len(cam_idx)
    nvar = 9 * ncams + 3 * npnts
    nt = nthreads or max_threads()
    cam_idx, pnt_idx, pt2d, x = _i64(cam_idx), _i64(pnt_idx), _f64(pt2d), _f64(x)
    if dense == "auto":
        dense = "scipy" if ncams > 400 else "c"
    if dense == "c":
        delta = np.empty(nvar)
        dr2 = C.c_double()
        jtr = np.empty(nvar) if want_jtr else None
        rc = lib().bao_lm_step_schur(cam_idx, pnt_idx, pt2d, ncams, npnts, nobs, x, float(lam), nt, delta,
                                     C.byref(dr2), None if jtr is None else jtr.ctypes.data_as(C.c_void_p))
        if rc:
            raise ArithmeticError("non-positive pivot in the reduced camera system")
        return (delta, dr2.value, jtr) if want_jtr else (delta, dr2.value)
    import scipy.linalg as sla
    r = cons(cam_idx, pnt_idx, pt2d, x, npnts, nt)
    vals = jac_coord(cam_idx, pnt_idx, x, npnts, nt)
    rows, cols = jac_structure(cam_idx, pnt_idx, npnts)
    jtr = mul_sparse(cols, rows, vals, r, nvar)
    del rows, cols
    n9 = 9 * ncams
    S, b = np.empty((n9, n9)), np.empty(n9)
    Vinv, hp = np.empty(9 * npnts), np.empty(3 * npnts)
    lib().bao_schur_system(cam_idx, pnt_idx, vals, jtr, ncams, npnts, nobs, float(lam), nt, S.reshape(-1), b,
                           Vinv, hp)
    cf = sla.cho_factor(S, lower=True, overwrite_a=True, check_finite=False)
    dc = sla.cho_solve(cf, b, check_finite=False)
    del S, cf
    delta = np.empty(nvar)
    lib().bao_schur_backsub(cam_idx, pnt_idx, vals, ncams, npnts, nobs, Vinv, hp, _f64(dc), nt, delta)
    v = vals.reshape(nobs, 2, 12)
    Jd = np.einsum("kij,kj->ki", v[:, :, :3], delta[: 3 * npnts].reshape(-1, 3)[pnt_idx - 1]) + \
        np.einsum("kij,kj->ki", v[:, :, 3:], delta[3 * npnts:].reshape(-1, 9)[cam_idx - 1])
    dr = Jd.reshape(-1) + r
    dr2 = float(np.linalg.norm(dr)) ** 2 / 2
    return (delta, dr2, jtr) if want_jtr else (delta, dr2)


def lm_solve(cam_idx, pnt_idx, pt2d, ncams, npnts, x0, params: LMParams | None = None, log_cap=512,
             solver="ldl") -> LMResult:
    """Levenberg_Marquardt(model, :LDL, <order>, :None, linesearch) (src/lm.jl:15-418).  solver="ldl": LDL' of the
    augmented matrix in natural order (src/ldl_aux.jl; small problems only); "schur": the same system in the
    elimination order of a fill-reducing permutation (points first, dense Cholesky of the camera system)."""
    nobs = len(cam_idx)
    p = params or default_params()
    st = LMStats()
    rows = (LMRow * log_cap)()
    x = _f64(x0).copy()
    lib().bao_lm_solve_ex(_i64(cam_idx), _i64(pnt_idx), _f64(pt2d), ncams, npnts, nobs, x, C.byref(p), C.byref(st),
                          C.cast(rows, C.c_void_p), log_cap, {"ldl": 0, "schur": 1}[solver])
    log = [dict(iter=r.iter, f=r.f, df=r.df, dfeas=r.dfeas, lam=r.lam, delta_norm=r.delta_norm, rho=r.rho,
                accepted=bool(r.accepted), acc_str=bool(r.acc_str)) for r in rows[: st.nrows]]
    return LMResult(STATUS[st.status], st.iter, st.objective, st.dual_feas, st.lambda_final, x, log, st.ldl_nnz)


def max_threads() -> int:
    return lib().bao_max_threads()
