#!/usr/bin/env python
"""Development probe (not product) of the mixed-precision factor: FP32 tensor-core Cholesky against LAPACK's FP32 one,
timings at the BASELINE sizes, the mixed LM step against the exact one.

    python scripts/mixed_dev.py chol 300 2304 8192 16128 | step | lm venice-1778 6
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bundleadjustment.jl_b200 as ba  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "chol"
if mode == "chol":
    for n in [int(a) for a in sys.argv[2:]] or [300, 2304]:
        rng = np.random.default_rng(n)
        if n <= 4096:  # spectrum 1 .. 1e4
            Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
            A = (Q * np.logspace(0, 4, n)) @ Q.T
            A = 0.5 * (A + A.T)
        else:
            A = rng.random((n, n))
            A = A + A.T
            A[np.arange(n), np.arange(n)] += n
        b = rng.normal(size=n)
        out = dict(n=n)
        for fp32 in (True, False):
            best = None
            for rep in range(3 if n > 4096 else 1):
                x, L, f, s = ba.lm.dbg_chol(A, b, want_L=(n <= 4096), fp32=fp32)
                best = (f, s) if best is None or f < best[0] else best
            tag = "fp32" if fp32 else "fp64"
            out[tag + "_factor_ms"] = best[0]
            out[tag + "_solve_ms"] = best[1]
            out[tag + "_tflops"] = n ** 3 / 3 / (best[0] * 1e-3) / 1e12
            out[tag + "_residual"] = float(np.linalg.norm(A @ x - b) / np.linalg.norm(b))
            if L is not None:
                L = np.tril(L)
                out[tag + "_LLt_err"] = float(np.linalg.norm(L @ L.T - A) / np.linalg.norm(A))
                if fp32:
                    Lr = np.linalg.cholesky(A.astype(np.float32)).astype(np.float64)
                    out["fp32_vs_lapack_spotrf"] = float(np.linalg.norm(L - Lr) / np.linalg.norm(Lr))
                    out["lapack_spotrf_LLt_err"] = float(np.linalg.norm(Lr @ Lr.T - A) / np.linalg.norm(A))
        print(json.dumps(out), flush=True)
elif mode == "step":
    for shape, lam in (((160, 10000, 50000), 30.0), ((160, 10000, 50000), 1e-3), ("trafalgar-257", 100.0),
                       ("trafalgar-257", 1e-2)):
        p = ba.synth.make_problem(shape)
        m = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs)
        m.set_solver("exact")
        d1, dr1, _, _, it1 = ba.lm_step(m, p.x0, lam)
        m.set_solver("mixed")
        d2, dr2, _, _, it2 = ba.lm_step(m, p.x0, lam)
        info = ba.lm.last_solve_info(m)
        d3, dr3, _, _, _ = ba.lm_step(m, p.x0, lam)
        m.close()
        print(json.dumps(dict(shape=str(shape), lam=lam, rel=float(np.linalg.norm(d2 - d1) / np.linalg.norm(d1)),
                              dr_rel=abs(dr2 - dr1) / dr1, cg_iters=int(it2), info=info,
                              rerun_identical=bool(np.array_equal(d2, d3)))), flush=True)
else:
    workload = sys.argv[2] if len(sys.argv) > 2 else "venice-1778"
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    p = ba.synth.make_problem(workload)
    m = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs)
    for solver in ("exact", "mixed", "mixed"):
        t0 = time.perf_counter()
        st = ba.Levenberg_Marquardt(m, "LDL", "AMD", "None", False, ite_max=iters - 1, solver=solver)
        dt = time.perf_counter() - t0
        print(json.dumps(dict(workload=workload, solver=solver, wall_s=dt, iters=st.iter, it_per_s=st.iter / dt,
                              objective=st.objective, status=st.status, cg=st.pcg_iters, timings_ms=st.timings_ms,
                              fallbacks=st.mixed_fallbacks, worst_rel=st.worst_solve_rel,
                              rows=[(r["f"], r["lam"], r["accepted"], r["pcg_iters"], r["solver"], r["solve_rel"])
                                    for r in st.rows])), flush=True)
    m.close()
