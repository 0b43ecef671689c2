"""One Levenberg_Marquardt() run on a synthetic workload (for ncu launch lists / traces of the LM loop).
usage: python scripts/lm_once.py <workload> [lm_iters] [pcg_max_iter]"""
import sys, time
sys.path.insert(0, ".")
import bundleadjustment.jl_b200 as ba
shape = sys.argv[1] if len(sys.argv) > 1 else "venice-1778"
shape = eval(shape) if shape.startswith("(") else shape
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 4
maxit = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
p = ba.synth.make_problem(shape)
m = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs)
t0 = time.time()
st = ba.Levenberg_Marquardt(m, "LDL", "AMD", "None", False, ite_max=iters - 1, pcg_max_iter=maxit)
print("status", st.status, "iters", st.iter, "pcg", [r["pcg_iters"] for r in st.rows], "time %.3fs" % (time.time() - t0),
      "objective %.9e" % st.objective, st.timings_ms)
