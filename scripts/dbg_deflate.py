"""GPU check of the PCG deflation feature: LM runs with deflation off / on for one workload, PCG iterations per
LM iteration, timings, final objective; then one damped solve repeated on the same system (harvest, then
deflated) compared with the undeflated step.  A failing configuration is re-run in a child process with
BAGPU_DEBUG_SYNC=1 so that the faulting launch is named.

usage: python scripts/dbg_deflate.py <workload|(ncams,npnts,nobs)> [lm_iters] [pcg_max_iter]"""
import os, subprocess, sys, time
sys.path.insert(0, ".")
import numpy as np
import bundleadjustment.jl_b200 as ba

shape = sys.argv[1] if len(sys.argv) > 1 else "venice-1778"
shape = eval(shape) if shape.startswith("(") else shape
lm_iters = int(sys.argv[2]) if len(sys.argv) > 2 else 8
maxit = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
child = os.environ.get("DBG_DEFLATE_CHILD")
p = ba.synth.make_problem(shape)
print("workload", shape, "ncams", p.ncams, "nobs", p.nobs, flush=True)


def run(k):
    m = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs)
    m.set_deflation(k)
    t0 = time.time()
    st = ba.Levenberg_Marquardt(m, "LDL", "AMD", "None", False, ite_max=lm_iters - 1, pcg_max_iter=maxit)
    dt = time.time() - t0
    print("deflate", k, "C call %.3fs" % st.elapsed_time, "status", st.status, "LM iters", st.iter, "pcg per iter", [r["pcg_iters"] for r in st.rows],
          "total", st.pcg_iters, "time %.3fs" % dt, "it/s %.2f" % (st.iter / dt), "objective %.12e" % st.objective,
          "timings", {a: round(b, 1) for a, b in st.timings_ms.items()}, flush=True)
    # repeated damped solve on one system
    outs = []
    for rep in range(3):
        t0 = time.time()
        d, dr2, obj, _, it = ba.lm_step(m, p.x0, 30.0, pcg_tol=1e-13, pcg_max_iter=maxit)
        outs.append((d, it, time.time() - t0))
    m.close()
    return st, outs


results = {}
for k in ([int(child)] if child else [0, 32]):
    try:
        results[k] = run(k)
    except Exception as e:  # noqa: BLE001
        print("deflate", k, "FAILED:", repr(e), flush=True)
        if not child:
            env = dict(os.environ, DBG_DEFLATE_CHILD=str(k), BAGPU_DEBUG_SYNC="1")
            r = subprocess.run([sys.executable] + sys.argv, env=env, capture_output=True, text=True, timeout=600)
            print("---- child with BAGPU_DEBUG_SYNC=1 ----\n" + r.stdout[-3000:] + r.stderr[-3000:], flush=True)
if 0 in results and 32 in results:
    s0, o0 = results[0]
    s1, o1 = results[32]
    print("objective rel diff %.3e" % (abs(s1.objective - s0.objective) / abs(s0.objective)))
    print("solution rel diff %.3e" % (np.linalg.norm(s1.solution - s0.solution) / np.linalg.norm(s0.solution)))
    ref = o0[0][0]
    for tag, outs in (("off", o0), ("on", o1)):
        print("lm_step deflate", tag, [(it, "%.3fs" % dt, "%.2e" % (np.linalg.norm(d - ref) / np.linalg.norm(ref)))
                                       for d, it, dt in outs])
