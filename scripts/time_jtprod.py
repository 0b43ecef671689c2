import sys, time; sys.path.insert(0, '.')
import ctypes as C
import numpy as np, torch
import bundleadjustment.jl_b200 as ba
p = ba.synth.make_problem("venice-1778")
m = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs)
L = ba._lib.lib(); h = m.handle
x = torch.from_numpy(p.x0).cuda(); v = torch.randn(2 * p.nobs, dtype=torch.float64, device="cuda")
out = torch.empty(p.nvar, dtype=torch.float64, device="cuda"); jv = torch.empty(2 * p.nobs, dtype=torch.float64, device="cuda")
vv = torch.randn(p.nvar, dtype=torch.float64, device="cuda")
for name, fn in (("jtprod", lambda: L.ba_jtprod_dev(h, C.c_void_p(x.data_ptr()), C.c_void_p(v.data_ptr()), C.c_void_p(out.data_ptr()))),
                 ("jprod", lambda: L.ba_jprod_dev(h, C.c_void_p(x.data_ptr()), C.c_void_p(vv.data_ptr()), C.c_void_p(jv.data_ptr())))):
    for _ in range(3): fn()
    L.ba_sync(h); t0 = time.perf_counter()
    for _ in range(20): fn()
    L.ba_sync(h); dt = (time.perf_counter() - t0) / 20
    print(name, "%.3f ms" % (dt * 1e3), "%.0f Mobs/s" % (p.nobs / dt / 1e6))
a = out.clone(); fn = None
L.ba_jtprod_dev(h, C.c_void_p(x.data_ptr()), C.c_void_p(v.data_ptr()), C.c_void_p(out.data_ptr())); L.ba_sync(h)
print("camera part bit-identical across runs:", bool(torch.equal(a[3 * p.npnts:], out[3 * p.npnts:])))
