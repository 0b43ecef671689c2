#!/usr/bin/env python
"""Development probe (not product): one LM run of a synthetic BAL-shaped problem with per-iteration phase times.

    BAGPU_TRACE=1 python scripts/lm_probe.py venice-1778 exact 6
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bundleadjustment.jl_b200 as ba  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "venice-1778"
solver = sys.argv[2] if len(sys.argv) > 2 else "auto"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 6
p = ba.synth.make_problem(workload)
m = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs)
for rep in range(2):
    t0 = time.perf_counter()
    st = ba.Levenberg_Marquardt(m, "LDL", "AMD", "None", False, ite_max=iters - 1, solver=solver, pcg_max_iter=4000)
    dt = time.perf_counter() - t0
    print(json.dumps(dict(workload=workload, solver=solver, rep=rep, wall_s=dt, iters=st.iter, it_per_s=st.iter / dt,
                          objective=st.objective, status=st.status, pcg_iters=st.pcg_iters, timings_ms=st.timings_ms,
                          capped=st.capped_solves, worst_rel=st.worst_solve_rel,
                          rows=[(r["f"], r["lam"], r["accepted"], r["pcg_iters"], r["solve_rel"]) for r in st.rows])))
m.close()
