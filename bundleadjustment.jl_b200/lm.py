"""``Levenberg_Marquardt`` entry point (src/lm.jl:15-418), device-resident.

Same positional arguments and keywords as the reference.  ``facto``/``perm``/``normalize`` are accepted
for signature compatibility: every combination of the reference solves the same system
``(J'J + lambda I) delta = -J'r``, which libbagpu solves on the GPU by a Schur
complement onto the camera system followed by block-Jacobi PCG.  The whole loop runs inside
``ba_lm_solve``; this wrapper only marshals parameters and results.
"""
from __future__ import annotations

import ctypes as C
import logging
import time
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from .model import BALNLPModel, FeasibilityResidual

STATUS = {0: "unknown", 1: "small_step", 2: "first_order", 3: "small_residual", 4: "acceptable",
          5: "neg_pred", 6: "exception", 7: "max_iter"}
log = logging.getLogger("bundleadjustment.lm")


@dataclass
class GenericExecutionStats:
    """The fields of SolverTools' GenericExecutionStats the reference fills (src/lm.jl:409-415)."""
    status: str
    solution: np.ndarray
    objective: float
    iter: int
    elapsed_time: float
    dual_feas: float
    # extras of the GPU path
    rows: list = field(default_factory=list)        # per-iteration log_row values (src/lm.jl:304)
    pcg_iters: int = 0
    lambda_final: float = 0.0
    timings_ms: dict = field(default_factory=dict)
    capped_solves: int = 0          # damped solves that stopped at pcg_max_iter (their steps are inexact)
    worst_solve_rel: float = 0.0    # largest relative residual a damped solve stopped at
    chol_n: int = 0                 # exact solver: order of the dense factorisations (0: PCG) and how many ran
    chol_count: int = 0
    mixed_fallbacks: int = 0        # mixed solver: damped solves that needed the FP64 factorisation after all


def default_params(**kw) -> _lib.LMParams:
    p = _lib.LMParams()
    _lib.lib().ba_lm_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError("unknown LM parameter %r" % k)
        setattr(p, k, v)
    return p


def Levenberg_Marquardt(model, facto="LDL", perm="AMD", normalize="None", linesearch=False, *, x=None,
                        restol=None, satol=None, srtol=None, oatol=None, ortol=None, atol=None, rtol=None,
                        nu_d=3.0, nu_m=3.0, lam=30.0, delta_d=2.0, ite_max=200, max_time=3600,
                        pcg_tol=None, pcg_max_iter=None, solver=None, verbose=False) -> GenericExecutionStats:
    """Levenberg_Marquardt(model, facto, perm, normalize, linesearch; x, tolerances, νd, νm, λ, δd, ite_max).

    ``model`` is a ``FeasibilityResidual`` (as in src/main.jl:27-30) or the ``BALNLPModel`` itself.
    ``max_time`` is accepted and, like in the reference, never consulted inside the loop
    (src/lm.jl:33,382: elapsed_time is only set after the loop).
    """
    if facto not in ("LDL", "QR") or perm not in ("AMD", "Metis") or normalize not in ("None", "A", "J"):
        raise ValueError("facto in {LDL,QR}, perm in {AMD,Metis}, normalize in {None,A,J} (src/lm.jl:15-19)")
    nlp = model.nlp if isinstance(model, FeasibilityResidual) else model
    if not isinstance(nlp, BALNLPModel):
        raise TypeError("model must be a BALNLPModel or its FeasibilityResidual")
    p = default_params(nu_d=nu_d, nu_m=nu_m, lam=lam, delta_d=delta_d, ite_max=int(ite_max),
                       linesearch=1 if linesearch else 0)
    if solver is not None:  # "auto" | "pcg" | "exact": how the damped system is solved (include/bagpu.h BA_SOLVER_*)
        p.solver = _lib.SOLVERS[solver]
    for k, v in dict(restol=restol, satol=satol, srtol=srtol, oatol=oatol, ortol=ortol, atol=atol, rtol=rtol,
                     pcg_tol=pcg_tol, pcg_max_iter=pcg_max_iter).items():
        if v is not None:
            setattr(p, k, v)
    xs = np.array(nlp.meta.x0 if x is None else x, dtype=np.float64, order="C", copy=True)
    if xs.size != nlp.meta.nvar:
        raise ValueError("x must have nvar = %d entries" % nlp.meta.nvar)
    rows = []

    def _cb(rowp, _user):
        r = rowp.contents
        d = dict(iter=r.iter, f=r.f, df=r.df, dfeas=r.dfeas, lam=r.lam, delta_norm=r.delta_norm, rho=r.rho,
                 accepted=bool(r.accepted), acc_str=bool(r.acc_str), pcg_iters=r.pcg_iters, ntimes=r.ntimes,
                 solver=_lib.SOLVER_NAMES.get(r.solver, "?"), converged=bool(r.converged), solve_rel=r.solve_rel)
        rows.append(d)
        if verbose:  # the 8 columns of log_row (src/lm.jl:120-121,304)
            print("%5d  %9.2e  %9.2e  %9.2e  %9.2e  %9.2e  %9.2e  %s" % (
                r.iter, r.f, r.df, r.dfeas, r.lam, r.delta_norm, r.rho, "acc" if r.acc_str else "rej"))

    cb = _lib.ITER_CB(_cb)
    st = _lib.LMStats()
    t0 = time.time()
    rc = _lib.lib().ba_lm_solve(nlp.handle, xs.ctypes.data_as(C.c_void_p), C.byref(p), C.byref(st),
                                C.cast(cb, C.c_void_p), None)
    _lib.check(rc, nlp.handle)
    elapsed = time.time() - t0
    nlp.counters.neval_cons += 1 + len(rows)
    nlp.counters.neval_jac += 2 + sum(1 for r in rows if r["accepted"])
    return GenericExecutionStats(
        status=STATUS.get(st.status, "unknown"), solution=xs, objective=st.objective, iter=int(st.iter),
        elapsed_time=elapsed, dual_feas=st.dual_feas, rows=rows, pcg_iters=int(st.pcg_iters_total),
        lambda_final=st.lambda_final,
        timings_ms=dict(eval=st.t_eval_ms, assemble=st.t_assemble_ms, pcg=st.t_pcg_ms, backsub=st.t_backsub_ms,
                        device_total=st.elapsed_s * 1e3, prepare=st.t_prepare_ms, schur_assembly=st.t_schur_ms,
                        cholesky=st.t_chol_ms),
        chol_n=int(st.chol_n), chol_count=int(st.chol_count), mixed_fallbacks=int(st.mixed_fallbacks),
        capped_solves=int(st.capped_solves), worst_solve_rel=st.worst_solve_rel)


def lm_step(nlp: BALNLPModel, x, lam, pcg_tol=1e-13, pcg_max_iter=500, want_jtr=False):
    """One damped solve (J'J + lam I) delta = -J'r at x: what ldl_factorize + ldl_solve! (or myqr +
    solve_qr!) deliver inside the loop (src/lm.jl:138-152,175-229).  Returns (delta, dr2, obj, jtr, iters)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    delta = np.empty(nlp.meta.nvar)
    jtr = np.empty(nlp.meta.nvar) if want_jtr else None
    dr2, obj, it = C.c_double(), C.c_double(), C.c_int32()
    rc = _lib.lib().ba_lm_step(nlp.handle, x.ctypes.data_as(C.c_void_p), float(lam), float(pcg_tol),
                               int(pcg_max_iter), delta.ctypes.data_as(C.c_void_p), C.byref(dr2), C.byref(obj),
                               None if jtr is None else jtr.ctypes.data_as(C.c_void_p), C.byref(it))
    _lib.check(rc, nlp.handle)
    return delta, dr2.value, obj.value, jtr, it.value


def last_solve_info(nlp: BALNLPModel) -> dict:
    """Outcome of the last damped solve on this model (ba_last_solve_info): which solver ran, whether it
    converged, the relative residual it stopped at, its iteration count."""
    sv, cv, it, rel = C.c_int32(), C.c_int32(), C.c_int32(), C.c_double()
    _lib.check(_lib.lib().ba_last_solve_info(nlp.handle, C.byref(sv), C.byref(cv), C.byref(rel), C.byref(it)))
    return dict(solver=_lib.SOLVER_NAMES.get(sv.value, "?"), converged=bool(cv.value), rel=rel.value, iters=it.value)


def dbg_chol(A, b, device=0, want_L=False, fp32=False):
    """Factor and solve a dense SPD system with the library's device Cholesky (ba_dbg_chol; fp32=True: the
    mixed-precision factor, ba_dbg_chol32).  Returns (x, L or None, factor_ms, solve_ms)."""
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    n = A.shape[0]
    x = np.empty(n)
    Lo = np.empty((n, n)) if want_L else None
    f, s = C.c_float(), C.c_float()
    fn = _lib.lib().ba_dbg_chol32 if fp32 else _lib.lib().ba_dbg_chol
    rc = fn(device, n, A.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p),
            x.ctypes.data_as(C.c_void_p), None if Lo is None else Lo.ctypes.data_as(C.c_void_p),
            C.byref(f), C.byref(s))
    if rc:
        raise _lib.BAError(rc, "ba_dbg_chol")
    return x, Lo, f.value, s.value
