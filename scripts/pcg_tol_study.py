#!/usr/bin/env python
"""How loose may pcg_tol be?  LM trajectories of the PCG solver at several tolerances against the exact solver
(explicit reduced camera system + Cholesky) on one BASELINE.json shape.   python scripts/pcg_tol_study.py venice-1778"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bundleadjustment.jl_b200 as ba  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "venice-1778"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
p = ba.synth.make_problem(workload)
m = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs)
ref = ba.Levenberg_Marquardt(m, "LDL", "AMD", "None", False, ite_max=iters - 1, solver="exact")
print(json.dumps(dict(workload=workload, solver="exact", iters=ref.iter, objective=ref.objective,
                      rows=[(r["f"], r["lam"], r["accepted"]) for r in ref.rows])), flush=True)
for tol in (1e-13, 1e-11, 1e-9, 1e-7, 1e-6, 1e-5, 1e-4):
    m.set_solver("pcg")   # releases the LM state: every tolerance starts without deflation vectors
    t0 = time.perf_counter()
    st = ba.Levenberg_Marquardt(m, "LDL", "AMD", "None", False, ite_max=iters - 1, solver="pcg", pcg_tol=tol,
                                pcg_max_iter=4000)
    dt = time.perf_counter() - t0
    dev = dict(f=0.0, lam=0.0)
    same = st.iter == ref.iter and [r["accepted"] for r in st.rows] == [r["accepted"] for r in ref.rows]
    for a, b in zip(st.rows, ref.rows):
        for k in dev:
            dev[k] = max(dev[k], abs(a[k] - b[k]) / abs(b[k]))
    print(json.dumps(dict(pcg_tol=tol, pcg_iters=st.pcg_iters, seconds=dt, same_decisions=same, max_rel_dev_f=dev["f"],
                          max_rel_dev_lambda=dev["lam"], objective_rel_dev=abs(st.objective - ref.objective) / ref.objective,
                          status=st.status, capped=st.capped_solves)), flush=True)
m.close()
