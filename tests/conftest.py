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


def rel_errors(a, b):
    """(norm-wise, floor-wise as in assert_rel, worst entry-wise) relative errors of a against b over the finite
    entries.  The entry-wise figure has no floor: entries that are small through cancellation show up there."""
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    fin = np.isfinite(b) & np.isfinite(a)
    d, bb = np.abs(a[fin] - b[fin]), np.abs(b[fin])
    nrm = float(np.linalg.norm(a[fin] - b[fin]) / max(np.linalg.norm(b[fin]), 1e-300))
    floor = float((d / (bb + bb.max() + 1e-300)).max()) if d.size else 0.0
    nz = bb > 0
    entry = float((d[nz] / bb[nz]).max()) if nz.any() else 0.0
    return nrm, floor, entry


def parity_report(name, **numbers):
    """Print the achieved errors of a parity test (pytest -rP / -s shows them) and append them to
    gpurun_out/parity_report.jsonl when that directory exists (copied to profiles/ for the record)."""
    import json
    line = json.dumps(dict(test=name, **numbers))
    print("[parity] " + line)
    d = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(d):
        with open(os.path.join(d, "parity_report.jsonl"), "a") as f:
            f.write(line + "\n")


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
