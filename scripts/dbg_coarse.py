import sys, time; sys.path.insert(0, '.')
import numpy as np
import bundleadjustment.jl_b200 as ba
shape = sys.argv[1]
shape = eval(shape) if shape.startswith("(") else shape
p = ba.synth.make_problem(shape)
m = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs)
ref = None
for n in [int(t) for t in sys.argv[2].split(",")]:
    m.set_coarse_clusters(n)
    for lam in (1e3, 30.0):
        t0 = time.time()
        d, dr2, obj, jtr, it = ba.lm_step(m, p.x0, lam, pcg_tol=1e-13, pcg_max_iter=5000)
        dt = time.time() - t0
        key = lam
        if ref is None or key not in ref:
            ref = ref or {}
            ref[key] = d
        print("clusters", n, "lambda", lam, "pcg iters", it, "time %.3fs" % dt, "rel diff vs first %.2e" % (np.linalg.norm(d - ref[key]) / np.linalg.norm(ref[key])))
