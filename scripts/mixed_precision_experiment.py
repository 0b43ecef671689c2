#!/usr/bin/env python
"""Row f4 of SURVEY.md section 8 as an experiment (host side, numpy/scipy; no GPU needed).

The reference's mixed-precision mode factorises in a lower precision (`facto_type`, src/lm.jl:92-98,165-173) while
residual and Jacobian stay FP64.  Two device variants were candidates; both are simulated here on the reduced camera
system S delta_c = b of a BASELINE.json shape, assembled by the oracle (oracle.bao_schur_system) from FP64 and from
FP32-rounded Jacobian values:

 (A) FP32 storage of the J blocks for the PCG operator (halves the 216 B/obs stream of the point-major pass), FP64
     vectors and dots, FP64 outer residual correction (iterative refinement around an inner block-Jacobi PCG on S32);
 (B) an FP32 Cholesky factor of the Jacobi-scaled S as the preconditioner of FP64 CG (what an FP32 version of the
     exact solver of ba_chol.cu would give).

Prints, per shape and lambda: block-Jacobi PCG iterations in FP64 to 1e-13, the inner/outer counts of (A) needed for a
1e-10 step, and the CG iterations of (B).    python scripts/mixed_precision_experiment.py [shape ...]
"""
import json
import os
import sys

import numpy as np
import scipy.linalg as sla

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bundleadjustment.jl_b200.synth as synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def schur(p, vals, jtr, lam, nt):
    n9 = 9 * p.ncams
    S, b = np.empty((n9, n9)), np.empty(n9)
    Vinv, hp = np.empty(9 * p.npnts), np.empty(3 * p.npnts)
    O.lib().bao_schur_system(p.cam_idx, p.pnt_idx, vals, jtr, p.ncams, p.npnts, p.nobs, float(lam), nt, S.reshape(-1), b,
                             Vinv, hp)
    return S, b


def block_jacobi(S):
    n = S.shape[0] // 9
    inv = np.empty((n, 9, 9))
    for c in range(n):
        inv[c] = np.linalg.inv(S[9 * c:9 * c + 9, 9 * c:9 * c + 9])
    return lambda r: np.einsum("cij,cj->ci", inv, r.reshape(n, 9)).reshape(-1)


def pcg(A, b, M, tol, maxit=20000):
    x = np.zeros_like(b)
    r = b.copy()
    z = M(r)
    p = z.copy()
    rz = rz0 = float(r @ z)
    for it in range(1, maxit + 1):
        q = A @ p
        a = rz / float(p @ q)
        x += a * p
        r -= a * q
        z = M(r)
        rzn = float(r @ z)
        if np.sqrt(rzn / rz0) <= tol:
            return x, it
        p = z + (rzn / rz) * p
        rz = rzn
    return x, maxit


def run(shape, lam):
    p = synth.make_problem(shape)
    nt = O.max_threads()
    r = O.cons(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.npnts, nt)
    vals = O.jac_coord(p.cam_idx, p.pnt_idx, p.x0, p.npnts, nt)
    rows, cols = O.jac_structure(p.cam_idx, p.pnt_idx, p.npnts)
    jtr = O.mul_sparse(cols, rows, vals, r, p.nvar)
    S, b = schur(p, vals, jtr, lam, nt)
    S32, _ = schur(p, vals.astype(np.float32).astype(np.float64), jtr, lam, nt)   # operator built from FP32 blocks
    xref = sla.cho_solve(sla.cho_factor(S, lower=True), b)
    rel = lambda x: float(np.linalg.norm(x - xref) / np.linalg.norm(xref))
    out = dict(shape=str(shape), lam=lam, n=S.shape[0], cond=float(np.linalg.cond(S)))
    M = block_jacobi(S)
    x, it64 = pcg(S, b, M, 1e-13)
    out["fp64_pcg"] = dict(iters=it64, step_err=rel(x))
    # (A) refinement around an inner PCG on the FP32-block operator
    best = None
    for inner_tol in (1e-2, 1e-4, 1e-6):
        M32 = block_jacobi(S32)
        x = np.zeros_like(b)
        tot, outer = 0, 0
        while outer < 40:
            res = b - S @ x                       # FP64 residual (one FP64 product per outer step)
            if np.linalg.norm(res) <= 1e-14 * np.linalg.norm(b) or rel(x) <= 1e-10:
                break
            d, it = pcg(S32, res, M32, inner_tol)
            x += d
            tot += it
            outer += 1
        cand = dict(inner_tol=inner_tol, outer=outer, inner_total=tot, step_err=rel(x))
        if cand["step_err"] <= 1e-10 and (best is None or tot + outer < best["inner_total"] + best["outer"]):
            best = cand
        out.setdefault("A_fp32_blocks", []).append(cand)
    out["A_best"] = best
    # (B) FP32 Cholesky of the Jacobi-scaled matrix as preconditioner of FP64 CG
    d = np.sqrt(np.diag(S))
    St = (S / d[:, None] / d[None, :])
    L32 = sla.cholesky(St.astype(np.float32), lower=True)
    Mch = lambda v: sla.cho_solve((L32, True), (v / d).astype(np.float32)).astype(np.float64) / d
    x, itb = pcg(S, b, Mch, 1e-13)
    out["B_fp32_cholesky_precond"] = dict(cg_iters=itb, step_err=rel(x))
    return out


if __name__ == "__main__":
    shapes = sys.argv[1:] or ["ladybug-49", "trafalgar-257"]
    for sh in shapes:
        for lam in (30.0, 0.37):
            print(json.dumps(run(sh, lam)), flush=True)
