import sys; sys.path.insert(0, '.')
import numpy as np
import bundleadjustment.jl_b200 as ba
mode, shape = sys.argv[1], sys.argv[2]
shape = eval(shape) if shape.startswith("(") else shape
p = ba.synth.make_problem(shape)
m = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs)
try:
    if mode == "step":
        d, dr2, obj, jtr, it = ba.lm_step(m, p.x0, 30.0, pcg_tol=1e-13, pcg_max_iter=50)
        print(mode, shape, "ok", dr2, obj, it)
    else:
        st = ba.Levenberg_Marquardt(m, "LDL", "AMD", "None", False, ite_max=1, pcg_max_iter=50)
        print(mode, shape, "ok", st.objective, st.iter)
except Exception as e:
    print(mode, shape, "FAIL", str(e)[-120:])
