import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


TOL = 1e-10  # north_star: residuals, Jacobian values and the LM step within 1e-10 relative


def assert_rel(a, b, tol=TOL, scale=None, what=""):
    """|a-b| <= tol * (|b| + scale) entrywise; scale defaults to max|b| (norm-wise floor, so entries that
    are small only through cancellation are judged against the size of the terms that cancelled).
    NaN/Inf must sit in the same places with the same sign."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    fin = np.isfinite(b)
    assert np.array_equal(np.isnan(a), np.isnan(b)), what + ": NaN pattern differs"
    assert np.array_equal(a[~fin & ~np.isnan(b)], b[~fin & ~np.isnan(b)]), what + ": Inf pattern differs"
    if not fin.any():
        return 0.0
    s = np.abs(b[fin]).max() if scale is None else scale
    err = np.abs(a[fin] - b[fin]) / (np.abs(b[fin]) + s + 1e-300)
    worst = float(err.max()) if err.size else 0.0
    assert worst <= tol, "%s: relative error %.3e > %.1e" % (what, worst, tol)
    return worst


def assert_jac_rel(vals, ref, tol=TOL, what="jac"):
    """Jacobian values: judged per column slot (24 slots per observation), because columns differ by
    many orders of magnitude (f*rho^4 for k2 against ~1e3 for the rotation)."""
    v = np.asarray(vals).reshape(-1, 24)
    r = np.asarray(ref).reshape(-1, 24)
    worst = 0.0
    for j in range(24):
        worst = max(worst, assert_rel(v[:, j], r[:, j], tol, what="%s slot %d" % (what, j)))
    return worst


@pytest.fixture(scope="session")
def golden():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "reference_runtests.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def ba():
    """The package under test.  If libbagpu.so has not been built yet (fresh checkout), build it in-tree
    first (nvcc cross-compiles sm_100a without a GPU) -- building is not a fallback: there is none."""
    import importlib.util as u
    spec = u.spec_from_file_location("_ba_build", os.path.join(ROOT, "bundleadjustment.jl_b200", "build.py"))
    mod = u.module_from_spec(spec)
    spec.loader.exec_module(mod)
    if not os.path.exists(mod.LIB):
        mod.build_lib()
    import bundleadjustment.jl_b200 as pkg
    return pkg


def small_problem(ba, shape=(7, 60, 260), **kw):
    return ba.synth.make_problem(shape, **kw)
