"""GPU parity tests of the operator surface (cons!/jac_structure!/jac_coord!/jprod!/jtprod!),
called through the C ABI exactly as the Julia glue would, against the CPU oracle on the same inputs.

Bar: jac_structure! bit-exact; residuals and Jacobian values within 1e-10 relative (north_star)."""
import ctypes as C

import numpy as np
import pytest

from conftest import TOL, assert_jac_rel, assert_rel, small_problem

pytestmark = pytest.mark.gpu


def _model(ba, p, **kw):
    return ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs, **kw)


# ---- reference golden vector (test/runtests.jl:15-27) through the CUDA path -------------------------
def test_golden_residual_vector_on_gpu(ba, golden):
    g = golden["residuals"]
    m = ba.BALNLPModel(g["cam_idx"], g["pnt_idx"], g["pt2d"], g["x"], 5, 1, 5)
    cx = m.cons(np.array(g["x"]))
    # device sincos / FMA contraction differ from Julia's libm in the last bits: 1e-10 relative
    assert_rel(cx, np.array(g["true_residuals"]), TOL, scale=np.abs(g["pt2d"]).max(), what="golden residuals")
    fr = ba.FeasibilityResidual(m)
    assert np.array_equal(fr.residual(np.array(g["x"])), cx)
    assert fr.nls_meta.nequ == 10 and fr.nls_meta.nnzj == 120 and fr.meta.name.endswith("-feasres")


# ---- the reference's own Python model (src/SolverScipy.py:34-72; tests/golden/make_scipy_reference_golden.py) -------
def test_residuals_and_jacobian_match_the_references_python_model_on_gpu(ba):
    """The CUDA path against reference-held code directly (no oracle in between): residuals of the reference's `fun`,
    and its Jacobian differentiated numerically in 80-bit arithmetic (accurate to ~1e-12)."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_scipy_model.json")) as f:
        g = json.load(f)
    ncams, npnts, nobs = g["shape"]
    x = np.array(g["x"])
    m = ba.BALNLPModel(np.array(g["cam_idx"]), np.array(g["pnt_idx"]), np.array(g["pt2d"]), x, ncams, npnts, nobs)
    cx, vals = m.cons_jac_coord_(x)
    m.close()
    assert np.abs(cx - np.array(g["residuals"])).max() <= TOL * np.abs(g["pt2d"]).max()
    ref = np.array(g["jac_vals"]).reshape(-1, 2, 12)
    err = np.abs(vals.reshape(-1, 2, 12) - ref) / np.abs(ref).max(axis=2, keepdims=True)
    assert err.max() <= TOL, err.max()


@pytest.mark.parametrize("variant", ["plain", "stress", "big_rotations"])
def test_residual_and_jacobian_match_oracle(ba, oracle, variant):
    p = small_problem(ba, shape=(9, 300, 1500), stress=variant == "stress", big_rotations=variant == "big_rotations")
    m = _model(ba, p)
    cx_ref = oracle.cons(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.npnts)
    vals_ref = oracle.jac_coord(p.cam_idx, p.pnt_idx, p.x0, p.npnts)
    scale = np.abs(p.pt2d).max()  # residual = projection - pt2d: judged against the size of its terms
    assert_rel(m.cons(p.x0), cx_ref, TOL, scale=scale, what="cons!")
    assert_jac_rel(m.jac_coord(p.x0), vals_ref, TOL, what="jac_coord!")
    cx, vals = m.cons_jac_coord_(p.x0)
    assert_rel(cx, m.cons(p.x0), 1e-14, scale=scale)  # fused and separate kernels agree to rounding
    assert_jac_rel(vals, m.jac_coord(p.x0), 1e-14)   # (FMA contraction may differ between instantiations)
    assert m.counters.neval_cons == 3 and m.counters.neval_jac == 3


def test_jac_structure_bit_exact(ba, oracle):
    p = small_problem(ba, shape=(9, 300, 1500))
    m = _model(ba, p)
    rows, cols = m.jac_structure()
    r_ref, c_ref = oracle.jac_structure(p.cam_idx, p.pnt_idx, p.npnts)
    assert rows.dtype == np.int64 and cols.dtype == np.int64
    assert np.array_equal(rows, r_ref) and np.array_equal(cols, c_ref)


@pytest.mark.parametrize("nobs", [1, 31, 32, 33, 127, 129])
def test_ragged_sizes(ba, oracle, nobs):
    # warp / block tails: every size must write exactly its own outputs
    rng = np.random.default_rng(nobs)
    p = small_problem(ba, shape=(5, 70, 300))
    sel = np.sort(rng.choice(p.nobs, nobs, replace=False))
    cam, pnt = p.cam_idx[sel], p.pnt_idx[sel]
    pt = p.pt2d.reshape(-1, 2)[sel].ravel()
    m = ba.BALNLPModel(cam, pnt, pt, p.x0, p.ncams, p.npnts, nobs)
    cx = np.full(2 * nobs, 7.0)
    vals = np.full(24 * nobs, 7.0)
    m.cons_jac_coord_(p.x0, cx, vals)
    assert_rel(cx, oracle.cons(cam, pnt, pt, p.x0, p.npnts), TOL, scale=np.abs(pt).max())
    assert_jac_rel(vals, oracle.jac_coord(cam, pnt, p.x0, p.npnts), TOL)
    rows, cols = m.jac_structure()
    r_ref, c_ref = oracle.jac_structure(cam, pnt, p.npnts)
    assert np.array_equal(rows, r_ref) and np.array_equal(cols, c_ref)


def test_empty_problem(ba):
    m = ba.BALNLPModel(np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0), np.arange(12.0), 1, 1, 0)
    assert m.cons(np.arange(12.0)).size == 0
    assert m.jac_coord(np.arange(12.0)).size == 0
    r, c = m.jac_structure()
    assert r.size == 0 and c.size == 0
    assert np.array_equal(m.jtprod_(np.arange(12.0), np.zeros(0)), np.zeros(12))


def test_theta_zero_and_z_zero_quirks(ba, oracle):
    # NaN residuals stay, all 24 entries of the block become 0
    p = small_problem(ba)
    x = p.x0.copy()
    c = p.cam_idx[5] - 1
    x[3 * p.npnts + 9 * c: 3 * p.npnts + 9 * c + 3] = 0.0
    m = _model(ba, p)
    cx, vals = m.cons_jac_coord_(x)
    cx_ref = oracle.cons(p.cam_idx, p.pnt_idx, p.pt2d, x, p.npnts)
    vals_ref = oracle.jac_coord(p.cam_idx, p.pnt_idx, x, p.npnts)
    hit = p.cam_idx == c + 1
    assert np.all(np.isnan(cx.reshape(-1, 2)[hit])) and np.all(vals.reshape(-1, 24)[hit] == 0.0)
    assert_rel(cx, cx_ref, TOL, scale=np.abs(p.pt2d).max())
    assert_jac_rel(vals, vals_ref, TOL)
    # z == 0 exactly
    x1 = np.array([0.25, -0.5, 1.0, 0.0, 0.0, 0.5, 0.1, 0.2, -1.0, 1e-3, 1e-5, 500.0])
    m1 = ba.BALNLPModel([1], [1], [0.0, 0.0], x1, 1, 1, 1)
    cx1, vals1 = m1.cons_jac_coord_(x1)
    ref1 = oracle.cons(np.array([1]), np.array([1]), np.zeros(2), x1, 1)
    assert np.array_equal(np.isnan(cx1), np.isnan(ref1)) and np.array_equal(np.isinf(cx1), np.isinf(ref1))
    assert np.all(vals1 == 0.0)


def test_jprod_jtprod_match_mul_sparse(ba, oracle):
    p = small_problem(ba, shape=(9, 300, 1500), stress=True)
    m = _model(ba, p)
    rng = np.random.default_rng(11)
    rows, cols = oracle.jac_structure(p.cam_idx, p.pnt_idx, p.npnts)
    vals = oracle.jac_coord(p.cam_idx, p.pnt_idx, p.x0, p.npnts)
    v = rng.normal(size=p.nvar)
    w = rng.normal(size=2 * p.nobs)
    Jv_ref = oracle.mul_sparse(rows, cols, vals, v, 2 * p.nobs)       # src/lma_aux.jl:194-212
    Jtw_ref = oracle.mul_sparse(cols, rows, vals, w, p.nvar)          # src/lm.jl:57
    assert_rel(m.jprod_(p.x0, v), Jv_ref, TOL, what="jprod!")
    Jtw = m.jtprod_(p.x0, w)
    assert_rel(Jtw[: 3 * p.npnts], Jtw_ref[: 3 * p.npnts], TOL, what="jtprod! points")
    g = Jtw[3 * p.npnts:].reshape(-1, 9)
    g_ref = Jtw_ref[3 * p.npnts:].reshape(-1, 9)
    for j in range(9):  # camera columns differ by orders of magnitude
        assert_rel(g[:, j], g_ref[:, j], TOL, what="jtprod! camera col %d" % j)


def test_input_validation(ba):
    p = small_problem(ba)
    m = _model(ba, p)
    with pytest.raises(ValueError):
        m.cons(p.x0[:-1])
    with pytest.raises(ValueError):
        m.jac_coord_(p.x0, np.empty(3))
    m.close()
    with pytest.raises(RuntimeError):
        m.cons(p.x0)


def test_sharded_handles_cover_the_problem(ba, oracle):
    # observation sharding (no collective on the evaluation path): rank outputs are slices
    p = small_problem(ba, shape=(9, 300, 1500))
    cx_ref = oracle.cons(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.npnts)
    vals_ref = oracle.jac_coord(p.cam_idx, p.pnt_idx, p.x0, p.npnts)
    r_ref, c_ref = oracle.jac_structure(p.cam_idx, p.pnt_idx, p.npnts)
    nr = 3
    covered = 0
    for r in range(nr):
        m = _model(ba, p, rank=r, nranks=nr)
        o0, o1 = m.obs_range
        assert o0 == covered
        covered = o1
        cx, vals = m.cons_jac_coord_(p.x0)
        assert_rel(cx, cx_ref[2 * o0:2 * o1], TOL, scale=np.abs(p.pt2d).max())
        assert_jac_rel(vals, vals_ref[24 * o0:24 * o1], TOL)
        rows, cols = m.jac_structure()
        assert np.array_equal(rows, r_ref[24 * o0:24 * o1]) and np.array_equal(cols, c_ref[24 * o0:24 * o1])
    assert covered == p.nobs


# ---- full BASELINE size: size-independent properties (the oracle would take too long to be the
# ---- checker for everything here, so it checks a slice and algebraic identities check the rest) ----
def test_venice_shape_properties(ba, oracle):
    p = ba.synth.make_problem("venice-1778")
    m = _model(ba, p)
    cx, vals = m.cons_jac_coord_(p.x0)
    assert np.all(np.isfinite(cx)) and np.all(np.isfinite(vals))
    # (1) oracle on a slice (last 20k observations: exercises the tail of the grid)
    s = slice(p.nobs - 20000, p.nobs)
    cam, pnt, pt = p.cam_idx[s], p.pnt_idx[s], p.pt2d[2 * s.start:]
    assert_rel(cx[2 * s.start:], oracle.cons(cam, pnt, pt, p.x0, p.npnts), TOL, scale=np.abs(pt).max())
    assert_jac_rel(vals[24 * s.start:], oracle.jac_coord(cam, pnt, p.x0, p.npnts), TOL)
    # (2) structure: closed-form checksums over all 120M entries
    rows, cols = m.jac_structure()
    k = np.arange(1, p.nobs + 1, dtype=np.int64)
    assert int(rows.sum()) == int((12 * (2 * k - 1) + 12 * 2 * k).sum())
    ip, ic = 3 * (p.pnt_idx - 1), 3 * p.npnts + 9 * (p.cam_idx - 1)
    assert int(cols.sum()) == int((2 * (3 * ip + 6 + 9 * ic + 45)).sum())
    assert np.array_equal(cols.reshape(-1, 24)[::9973, :12], cols.reshape(-1, 24)[::9973, 12:])
    del rows, cols
    # (3) adjointness <J v, w> == <v, J' w> and jprod == vals contracted with v (linearity of the COO form)
    rng = np.random.default_rng(1)
    v = rng.normal(size=p.nvar)
    w = rng.normal(size=2 * p.nobs)
    Jv = m.jprod_(p.x0, v)
    Jtw = m.jtprod_(p.x0, w)
    lhs, rhs = float(Jv @ w), float(v @ Jtw)
    assert abs(lhs - rhs) <= 1e-11 * (np.linalg.norm(Jv) * np.linalg.norm(w))
    V = vals.reshape(-1, 2, 12)
    vp = v[: 3 * p.npnts].reshape(-1, 3)[p.pnt_idx - 1]
    vc = v[3 * p.npnts:].reshape(-1, 9)[p.cam_idx - 1]
    Jv2 = np.einsum("kij,kj->ki", V[:, :, :3], vp) + np.einsum("kij,kj->ki", V[:, :, 3:], vc)
    assert_rel(Jv, Jv2.ravel(), TOL, what="jprod vs vals")


def test_model_from_bal_file(ba, oracle, tmp_path):
    # BALNLPModel(filename) (src/BALNLPModels.jl:91-106) through the BAL reader (src/ReadFiles.jl:9-53)
    from bundleadjustment.jl_b200 import balio
    p = small_problem(ba)
    d = tmp_path / "Synth"
    d.mkdir()
    f = d / ("problem-%d-%d-pre.txt.bz2" % (p.ncams, p.npnts))
    balio.write_problem(str(f), p)
    m = ba.BALNLPModel.from_file(str(f))
    assert m.meta.name == "Synth-%d-%d" % (p.ncams, p.npnts)            # name(), src/BALNLPModels.jl:58-68
    assert (m.meta.nvar, m.meta.ncon, m.meta.nnzj) == (p.nvar, 2 * p.nobs, 24 * p.nobs)
    assert np.array_equal(m.meta.x0, p.x0)
    assert_rel(m.cons(m.meta.x0), oracle.cons(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.npnts), TOL,
               scale=np.abs(p.pt2d).max())


def test_profiling_switch_and_kernel_time(ba):
    import ctypes as C
    p = small_problem(ba, shape=(9, 300, 1500))
    m = _model(ba, p)
    L = ba._lib.lib()
    ms = C.c_float()
    m.cons_jac_coord_(p.x0)
    assert L.ba_last_eval_ms(m.handle, C.byref(ms)) == ba._lib.BA_ERR_ARG     # profiling is off by default
    assert b"ba_set_profiling" in L.ba_last_error(m.handle)
    ba._lib.check(L.ba_set_profiling(m.handle, 1), m.handle)
    cx, vals = m.cons_jac_coord_(p.x0)
    ba._lib.check(L.ba_last_eval_ms(m.handle, C.byref(ms)), m.handle)
    assert 0.0 < ms.value < 50.0
    ba._lib.check(L.ba_set_profiling(m.handle, 0), m.handle)
    cx2, vals2 = m.cons_jac_coord_(p.x0)
    assert np.array_equal(cx, cx2) and np.array_equal(vals, vals2)   # same kernel with and without the events


def test_jtprod_camera_part_is_reproducible(ba):
    # point-major problems: the camera side of J'v is an ordered camera-major sum (no FP64 atomics)
    p = small_problem(ba, shape=(9, 300, 1500))
    m = _model(ba, p)
    w = np.random.default_rng(3).normal(size=2 * p.nobs)
    a = m.jtprod_(p.x0, w)
    for _ in range(3):
        b = m.jtprod_(p.x0, w)
        assert np.array_equal(a[3 * p.npnts:], b[3 * p.npnts:])


@pytest.mark.gpu
def test_fp64_peak_probe_is_sane(ba):
    import ctypes as C
    v = C.c_double()
    assert ba._lib.lib().ba_measure_fp64_peak(0, C.byref(v)) == 0
    assert 5.0 < v.value < 200.0, v.value          # B200: a few tens of TFLOP/s
