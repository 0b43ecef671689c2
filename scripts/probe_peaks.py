#!/usr/bin/env python
"""Development probe: tensor / FMA rates an FP32-accurate factorisation could build on (ba_dbg_probe_peak);
profiles/r02_probe_peaks.txt.  python scripts/probe_peaks.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bundleadjustment.jl_b200 as ba  # noqa: E402

L = ba._lib.lib()
f = L.ba_dbg_probe_peak
for kind, name in ((0, "tf32 mma.sync m16n8k8"), (1, "bf16 mma.sync m16n8k16"), (2, "fp32 fma")):
    t = C.c_double()
    rc = f(0, kind, C.byref(t))
    print(name, rc, round(t.value, 1), "TFLOP/s")
