import ctypes as C, sys, os
sys.path.insert(0, os.getcwd())
import bundleadjustment.jl_b200 as ba
L = ba._lib.lib()
f = L.ba_dbg_probe_peak
f.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double)]
for kind, name in ((0, "tf32 mma.sync m16n8k8"), (1, "bf16 mma.sync m16n8k16"), (2, "fp32 fma")):
    t = C.c_double()
    rc = f(0, kind, C.byref(t))
    print(name, rc, round(t.value, 1), "TFLOP/s")
