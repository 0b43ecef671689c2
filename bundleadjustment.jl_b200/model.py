"""Host-side mirror of the reference's operator surface for the hot path.

The reference's host language is Julia (not available in this image), so the mirror above the C ABI
is Python: same names, argument meaning and error behaviour as

* ``BALNLPModel``            -- src/BALNLPModels.jl:79-106 (struct + ctor, meta fields)
* ``cons!/jac_structure!/jac_coord!`` -- src/BALNLPModels.jl:115-206 (Julia's ``!`` becomes a trailing ``_``)
* ``FeasibilityResidual``    -- NLPModels 0.12.4 adaptor used at src/main.jl:27
  (``residual!``, ``jac_structure_residual!``, ``jac_coord_residual!``, ``jprod_residual!``, ``jtprod_residual!``)

Every method is ONE call into libbagpu.so (ctypes here, ``ccall`` in julia/BALGPUModels.jl).
Indices are 1-based int64 and ``x = [points; cameras]`` exactly as in the reference.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(a, n=None, name="array"):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if n is not None and a.size != n:
        raise ValueError("%s: expected %d elements, got %d" % (name, n, a.size))
    return a


def _out(a, n, dtype, name):
    if a is None:
        return np.empty(n, dtype=dtype)
    if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags.c_contiguous or a.size != n:
        raise ValueError("%s must be a C-contiguous %s array of %d elements" % (name, np.dtype(dtype).name, n))
    return a


@dataclass
class NLPModelMeta:
    """The NLPModelMeta fields the reference sets (src/BALNLPModels.jl:102)."""
    nvar: int
    ncon: int
    x0: np.ndarray
    lcon: float = 0.0   # lcon = ucon = 0 for every constraint (kept scalar: 2*nobs zeros otherwise)
    ucon: float = 0.0
    nnzj: int = 0
    name: str = "BAL"


@dataclass
class Counters:
    neval_cons: int = 0
    neval_jac: int = 0
    neval_jprod: int = 0
    neval_jtprod: int = 0


def name(filename: str) -> str:
    """src/BALNLPModels.jl:58-68: 'LadyBug/problem-49-7776-pre.txt.bz2' -> 'LadyBug-49-7776'."""
    k = filename.index("/")
    l = k + 8
    while filename[l] != "p":
        l += 1
    return filename[:k] + filename[k + 8: l - 1]


class BALNLPModel:
    """GPU-backed ``BALNLPModel`` (src/BALNLPModels.jl:79-106).

    Build it from arrays (``cams_indices``/``pnts_indices`` 1-based, ``pt2d`` interleaved, ``x0``) --
    what ``readfile`` returns (src/ReadFiles.jl:9-53) -- or from a BAL file with ``from_file``.
    ``rank``/``nranks`` select the observation shard of a one-process-per-GPU run; ``ngpus`` (an int, or "all")
    puts several GPUs behind this one model in this one process (ba_create_multi): every method then behaves as
    on one GPU, with full-length arrays.
    """

    def __init__(self, cams_indices, pnts_indices, pt2d, x0, ncams, npnts, nobs=None, *, name="BAL",
                 device=0, rank=0, nranks=1, ngpus=None):
        self.cams_indices = np.ascontiguousarray(cams_indices, dtype=np.int64)
        self.pnts_indices = np.ascontiguousarray(pnts_indices, dtype=np.int64)
        self.nobs = int(len(self.cams_indices) if nobs is None else nobs)
        self.npnts, self.ncams = int(npnts), int(ncams)
        if self.cams_indices.size != self.nobs or self.pnts_indices.size != self.nobs:
            raise ValueError("index vectors must have nobs entries")
        self.pt2d = _f64(pt2d, 2 * self.nobs, "pt2d")
        nvar = 9 * self.ncams + 3 * self.npnts
        self.meta = NLPModelMeta(nvar=nvar, ncon=2 * self.nobs, x0=_f64(x0, nvar, "x0").copy(),
                                 nnzj=2 * self.nobs * 12, name=name)
        self.counters = Counters()
        self.device, self.rank, self.nranks = int(device), int(rank), int(nranks)
        L = _lib.lib()
        h = C.c_void_p()
        self.ngpus = None
        if ngpus is not None:
            if nranks != 1:
                raise ValueError("ngpus (one process, several GPUs) and rank/nranks (one process per GPU) exclude each other")
            n = 0 if ngpus == "all" else int(ngpus)
            rc = L.ba_create_multi(self.ncams, self.npnts, self.nobs, _ptr(self.cams_indices),
                                   _ptr(self.pnts_indices), _ptr(self.pt2d), n, None, C.byref(h))
            self.ngpus = ngpus
        elif nranks == 1:
            rc = L.ba_create(self.ncams, self.npnts, self.nobs, _ptr(self.cams_indices), _ptr(self.pnts_indices),
                             _ptr(self.pt2d), self.device, C.byref(h))
        else:
            rc = L.ba_create_sharded(self.ncams, self.npnts, self.nobs, _ptr(self.cams_indices),
                                     _ptr(self.pnts_indices), _ptr(self.pt2d), self.device, self.rank,
                                     self.nranks, C.byref(h))
        self._h = h
        if rc != 0:
            msg = L.ba_last_error(h).decode() if h else "ba_create failed"
            if h:
                L.ba_destroy(h)
            self._h = None
            raise _lib.BAError(rc, msg)
        r = [C.c_int64() for _ in range(4)]
        _lib.check(L.ba_shard_range(h, *[C.byref(v) for v in r]), h)
        self.obs_range = (r[0].value, r[1].value)   # 0-based half-open, this rank's observations
        self.pnt_range = (r[2].value, r[3].value)
        self.nobs_local = self.obs_range[1] - self.obs_range[0]

    @classmethod
    def from_file(cls, filename, **kw):
        from .balio import readfile
        cams, pnts, pt2d, x0, ncams, npnts, nobs = readfile(filename)
        nm = filename
        try:
            nm = name("/".join(filename.replace("\\", "/").split("/")[-2:]))
        except (ValueError, IndexError):
            pass
        return cls(cams, pnts, pt2d, x0, ncams, npnts, nobs, name=nm, **kw)

    # ---- lifetime ----------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().ba_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        if not self._h:
            raise RuntimeError("model is closed")
        return self._h

    def set_coarse_clusters(self, n: int):
        """PCG preconditioner of the LM solve: block-Jacobi plus an additive coarse level over ``n`` camera
        clusters (default 16, at most 24; 0 = plain block-Jacobi).  Changes iteration counts, not solutions."""
        _lib.check(_lib.lib().ba_set_coarse_clusters(self.handle, int(n)), self.handle)

    def set_solver(self, solver: str):
        """How the damped LM system is solved: "auto" (exact up to 2048 cameras, PCG above), "pcg" (matrix-free
        preconditioned CG), "exact" (explicit reduced camera system + dense FP64 Cholesky + refinement) or "mixed"
        (the same system factorised in FP32 on the tensor cores, preconditioning FP64 CG on the FP64 operator:
        the reference's facto_type < T mode, src/lm.jl:92-98)."""
        _lib.check(_lib.lib().ba_set_solver(self.handle, _lib.SOLVERS[solver]), self.handle)

    def set_deflation(self, k: int):
        """PCG of the LM solve: add up to ``k`` (<= 32; default 32; 0 = off) Ritz vectors harvested from the PCG
        solves themselves to the coarse level, plus up to 16 refreshed after every solve.  Changes iteration
        counts, not solutions."""
        _lib.check(_lib.lib().ba_set_deflation(self.handle, int(k)), self.handle)

    # ---- NLPModels surface ---------------------------------------------------------------------
    def obj(self, x):
        """NLPModels.obj: identically 0 (src/BALNLPModels.jl:109)."""
        return 0.0

    def grad_(self, x, g):
        """NLPModels.grad!: zeros (src/BALNLPModels.jl:112)."""
        g[...] = 0
        return g

    def cons_(self, x, cx=None):
        """NLPModels.cons!(nlp, x, cx): cx = projection - pt2d (src/BALNLPModels.jl:115-122)."""
        self.counters.neval_cons += 1
        x = _f64(x, self.meta.nvar, "x")
        cx = _out(cx, 2 * self.nobs_local, np.float64, "cx")
        _lib.check(_lib.lib().ba_residual(self.handle, _ptr(x), _ptr(cx)), self.handle)
        return cx

    def cons(self, x):
        return self.cons_(x)

    def jac_structure_(self, rows=None, cols=None):
        """NLPModels.jac_structure!(nlp, rows, cols): 1-based Int64 COO pattern (src/BALNLPModels.jl:125-158)."""
        self.counters.neval_jac += 1  # the reference bumps neval_jac here too (:126)
        n = 24 * self.nobs_local
        rows = _out(rows, n, np.int64, "rows")
        cols = _out(cols, n, np.int64, "cols")
        _lib.check(_lib.lib().ba_jac_structure(self.handle, _ptr(rows), _ptr(cols)), self.handle)
        return rows, cols

    def jac_structure(self):
        return self.jac_structure_()

    def jac_coord_(self, x, vals=None):
        """NLPModels.jac_coord!(nlp, x, vals) (src/BALNLPModels.jl:161-206)."""
        self.counters.neval_jac += 1
        x = _f64(x, self.meta.nvar, "x")
        vals = _out(vals, 24 * self.nobs_local, np.float64, "vals")
        _lib.check(_lib.lib().ba_jac_coord(self.handle, _ptr(x), _ptr(vals)), self.handle)
        return vals

    def jac_coord(self, x):
        return self.jac_coord_(x)

    def cons_jac_coord_(self, x, cx=None, vals=None):
        """cons! and jac_coord! in one pass over the observations (ba_residual_jac)."""
        self.counters.neval_cons += 1
        self.counters.neval_jac += 1
        x = _f64(x, self.meta.nvar, "x")
        cx = _out(cx, 2 * self.nobs_local, np.float64, "cx")
        vals = _out(vals, 24 * self.nobs_local, np.float64, "vals")
        _lib.check(_lib.lib().ba_residual_jac(self.handle, _ptr(x), _ptr(cx), _ptr(vals)), self.handle)
        return cx, vals

    def jprod_(self, x, v, Jv=None):
        """NLPModels.jprod!(nlp, x, v, Jv): Jv = J(x) v, matrix-free (semantics of
        mul_sparse!(rows, cols, vals, v, ...), src/lma_aux.jl:194-212)."""
        self.counters.neval_jprod += 1
        x = _f64(x, self.meta.nvar, "x")
        v = _f64(v, self.meta.nvar, "v")
        Jv = _out(Jv, 2 * self.nobs_local, np.float64, "Jv")
        _lib.check(_lib.lib().ba_jprod(self.handle, _ptr(x), _ptr(v), _ptr(Jv)), self.handle)
        return Jv

    def jtprod_(self, x, v, Jtv=None):
        """NLPModels.jtprod!(nlp, x, v, Jtv): Jtv = J(x)' v (mul_sparse! with rows/cols swapped,
        src/lm.jl:57,356,370).  ``v`` covers this rank's observations."""
        self.counters.neval_jtprod += 1
        x = _f64(x, self.meta.nvar, "x")
        v = _f64(v, 2 * self.nobs_local, "v")
        Jtv = _out(Jtv, self.meta.nvar, np.float64, "Jtv")
        _lib.check(_lib.lib().ba_jtprod(self.handle, _ptr(x), _ptr(v), _ptr(Jtv)), self.handle)
        return Jtv


@dataclass
class NLSMeta:
    nequ: int
    nnzj: int


class FeasibilityResidual:
    """NLPModels 0.12.4 ``FeasibilityResidual(nlp)`` (src/main.jl:27): the NLS view whose residual is
    ``cons(x) - lcon`` with ``lcon = 0``; pure forwarding, no arithmetic of its own."""

    def __init__(self, nlp: BALNLPModel):
        self.nlp = nlp
        self.meta = NLPModelMeta(nvar=nlp.meta.nvar, ncon=0, x0=nlp.meta.x0, nnzj=0, name=nlp.meta.name + "-feasres")
        self.nls_meta = NLSMeta(nequ=nlp.meta.ncon, nnzj=nlp.meta.nnzj)
        self.counters = nlp.counters

    def residual_(self, x, Fx=None):
        return self.nlp.cons_(x, Fx)

    def residual(self, x):
        return self.nlp.cons_(x)

    def jac_structure_residual_(self, rows=None, cols=None):
        return self.nlp.jac_structure_(rows, cols)

    def jac_coord_residual_(self, x, vals=None):
        return self.nlp.jac_coord_(x, vals)

    def jac_coord_residual(self, x):
        return self.nlp.jac_coord_(x)

    def jprod_residual_(self, x, v, Jv=None):
        return self.nlp.jprod_(x, v, Jv)

    def jtprod_residual_(self, x, v, Jtv=None):
        return self.nlp.jtprod_(x, v, Jtv)
