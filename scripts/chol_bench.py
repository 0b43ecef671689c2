#!/usr/bin/env python
"""Development probe: device Cholesky (ba_dbg_chol) on diagonally dominant random SPD matrices; prints factor / solve
times and TFLOP/s (n^3/3).  python scripts/chol_bench.py 2304 4096 8192 16128"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bundleadjustment.jl_b200 as ba  # noqa: E402

for n in [int(a) for a in sys.argv[1:]] or [2304, 4096, 8192]:
    rng = np.random.default_rng(n)
    A = rng.random((n, n), dtype=np.float64)
    A = A + A.T
    A[np.arange(n), np.arange(n)] += n
    b = rng.normal(size=n)
    best = None
    for rep in range(3):
        x, _, f, s = ba.lm.dbg_chol(A, b)
        best = (f, s) if best is None or f < best[0] else best
    res = float(np.linalg.norm(A @ x - b) / np.linalg.norm(b))
    print(json.dumps(dict(n=n, factor_ms=best[0], solve_ms=best[1], tflops=n ** 3 / 3 / (best[0] * 1e-3) / 1e12,
                          residual=res)), flush=True)
import ctypes as C  # noqa: E402
t = C.c_double()
L = ba._lib.lib()
L.ba_measure_fp64_peak(0, C.byref(t)); print(json.dumps(dict(fp64_fma_peak_tflops=t.value)))
L.ba_measure_fp64_mma_peak(0, C.byref(t)); print(json.dumps(dict(fp64_mma_peak_tflops=t.value)))
