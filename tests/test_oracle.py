"""CPU tests of the oracle (the checker): pinned against the reference's own golden vectors and
known-answer tests (test/runtests.jl), then anchored by finite differences and dense algebra where
the reference has no test of its own (jac_structure!, jac_coord!, LDL, LM loop)."""
import numpy as np
import pytest

from conftest import assert_rel, small_problem


# ---- reference test/runtests.jl:6-8 -----------------------------------------------------------
def test_rodrigues_known_answer_bit_exact(oracle, golden):
    g = golden["rodrigues"]
    out = oracle.rodrigues_rotation(g["r"], g["x"])
    assert out.tolist() == g["expect"]  # the reference asserts == (bit-exact)


def test_scaling_and_projection_known_answers(oracle, golden):
    g = golden["scaling_factor"]
    assert oracle.scaling_factor(g["point"], g["k1"], g["k2"]) == g["expect"]
    a = golden["projection"]["args_x_y_z_rx_ry_rz_tx_ty_tz_f_k1_k2"]
    out = oracle.projection_jump(a[0:3], a[3:6], a[6:9], a[9], a[10], a[11])
    # runtests.jl:8 asserts == [-7 -7]; the value carries the rounding of Rodrigues at r=(1,1,1)
    assert np.allclose(out, golden["projection"]["expect"], rtol=1e-14, atol=0)


# ---- reference test/runtests.jl:15-27: residuals! golden vector, norm(...) == 0 -------------------
def test_residual_golden_vector_bit_exact(oracle, golden):
    g = golden["residuals"]
    cx = oracle.cons(np.array(g["cam_idx"]), np.array(g["pnt_idx"]), np.array(g["pt2d"]), np.array(g["x"]),
                     g["npnts"])
    assert np.linalg.norm(np.array(g["true_residuals"]) - cx) == 0.0


def test_residual_threads_do_not_change_bits(oracle, ba):
    p = small_problem(ba)
    a = oracle.cons(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.npnts, nthreads=1)
    b = oracle.cons(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.npnts, nthreads=4)
    assert np.array_equal(a, b)


# ---- jac_structure! (src/BALNLPModels.jl:125-158): hand-checkable formula -------------------------
def test_jac_structure_formula(oracle, golden):
    g = golden["residuals"]
    rows, cols = oracle.jac_structure(np.array(g["cam_idx"]), np.array(g["pnt_idx"]), g["npnts"])
    # npnts = 1 => ip = 0, ic = 3 + 9(c-1)
    assert rows[:24].tolist() == [1] * 12 + [2] * 12
    assert cols[:24].tolist() == list(range(1, 13)) * 2
    assert rows[24:48].tolist() == [3] * 12 + [4] * 12
    assert cols[24:48].tolist() == ([1, 2, 3] + list(range(13, 22))) * 2
    assert rows.dtype == np.int64 and cols.dtype == np.int64


def test_jac_structure_generic(oracle, ba):
    p = small_problem(ba)
    rows, cols = oracle.jac_structure(p.cam_idx, p.pnt_idx, p.npnts)
    k = np.arange(p.nobs)
    assert np.array_equal(rows.reshape(-1, 24)[:, 0], 2 * k + 1)
    assert np.array_equal(rows.reshape(-1, 24)[:, 23], 2 * k + 2)
    c = cols.reshape(-1, 24)
    assert np.array_equal(c[:, 0], 3 * (p.pnt_idx - 1) + 1)
    assert np.array_equal(c[:, 3], 3 * p.npnts + 9 * (p.cam_idx - 1) + 1)
    assert np.array_equal(c[:, :12], c[:, 12:])
    assert cols.max() <= p.nvar and cols.min() >= 1


# ---- jac_coord! (src/BALNLPModels.jl:161-206): unpinned by the reference; finite differences ------
def _fd_jac(oracle, p, x, k, h=1e-6):
    """central differences of the two residuals of observation k wrt its 12 parameters"""
    cam, pnt = p.cam_idx[k:k + 1], p.pnt_idx[k:k + 1]
    pt = p.pt2d[2 * k:2 * k + 2]
    cols = list(range(3 * (pnt[0] - 1), 3 * (pnt[0] - 1) + 3)) + \
        list(range(3 * p.npnts + 9 * (cam[0] - 1), 3 * p.npnts + 9 * (cam[0] - 1) + 9))
    J = np.empty((2, 12))
    for j, c in enumerate(cols):
        step = h * max(1.0, abs(x[c]))
        xp, xm = x.copy(), x.copy()
        xp[c] += step
        xm[c] -= step
        J[:, j] = (oracle.cons(cam, pnt, pt, xp, p.npnts) - oracle.cons(cam, pnt, pt, xm, p.npnts)) / (2 * step)
    return J


@pytest.mark.parametrize("variant", ["plain", "stress", "big_rotations"])
def test_jac_coord_matches_finite_differences(oracle, ba, variant):
    kw = dict(stress=variant == "stress", big_rotations=variant == "big_rotations")
    p = small_problem(ba, **kw)
    vals = oracle.jac_coord(p.cam_idx, p.pnt_idx, p.x0, p.npnts).reshape(-1, 2, 12)
    for k in (0, 17, 101, p.nobs - 1):
        J = _fd_jac(oracle, p, p.x0, k)
        scale = np.abs(J).max(axis=0, keepdims=True) + 1e-300
        if variant != "stress":  # k1, k2 ~ 1e-7, 1e-13: relative steps underflow the FD there
            J, v, scale = J[:, :9], vals[k][:, :9], scale[:, :9]
        else:
            v = vals[k]
        assert np.all(np.abs(v - J) <= 5e-6 * scale), (variant, k)


def test_jac_coord_golden_point_fd(oracle, golden):
    g = golden["residuals"]
    cam, pnt, x = np.array(g["cam_idx"]), np.array(g["pnt_idx"]), np.array(g["x"])
    vals = oracle.jac_coord(cam, pnt, x, 1).reshape(-1, 2, 12)

    class P:  # minimal view for _fd_jac
        cam_idx, pnt_idx, pt2d, npnts = cam, pnt, np.array(g["pt2d"]), 1
    for k in range(5):
        J = _fd_jac(oracle, P, x, k)
        scale = np.abs(J).max(axis=0, keepdims=True)
        assert np.all(np.abs(vals[k][:, :9] - J[:, :9]) <= 5e-6 * scale[:, :9])


# ---- the reference's own Python model (src/SolverScipy.py:34-72), tests/golden/make_scipy_reference_golden.py ----
@pytest.fixture(scope="module")
def scipy_model_golden():
    import json
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden", "reference_scipy_model.json")) as f:
        g = json.load(f)
    return {k: (np.array(v) if isinstance(v, list) else v) for k, v in g.items()}


def test_residuals_match_the_references_python_model(oracle, scipy_model_golden):
    """cons! of the oracle against `fun` of the reference's Python implementation (an independent statement of the
    same camera model by the reference's authors) on 150 observations, 6 cameras with rotations up to 1.3 rad."""
    g = scipy_model_golden
    npnts = int(g["shape"][1])
    cx = oracle.cons(g["cam_idx"], g["pnt_idx"], g["pt2d"], g["x"], npnts)
    ref = g["residuals"]
    scale = np.abs(g["pt2d"]).max()       # the residual is a difference of pixel coordinates of this size
    assert np.abs(cx - ref).max() <= 1e-12 * scale


def test_hand_derived_jacobian_matches_the_references_python_model(oracle, scipy_model_golden):
    """jac_coord! (the hand-derived blocks of src/JacobianByHand.jl as restated by the oracle) against the Jacobian of
    the reference's Python `fun`, differentiated numerically in 80-bit arithmetic (accurate to ~1e-12): every one of
    the 24 entries of every observation, relative to the largest entry of its row."""
    g = scipy_model_golden
    npnts = int(g["shape"][1])
    vals = oracle.jac_coord(g["cam_idx"], g["pnt_idx"], g["x"], npnts).reshape(-1, 2, 12)
    ref = g["jac_vals"].reshape(-1, 2, 12)
    scale = np.abs(ref).max(axis=2, keepdims=True)
    err = np.abs(vals - ref) / scale
    assert err.max() <= 1e-10, err.max()
    # and entry-wise for the entries that are not cancellation residues (>= 1e-6 of the row's largest)
    big = np.abs(ref) >= 1e-6 * scale
    assert (np.abs(vals - ref)[big] / np.abs(ref)[big]).max() <= 1e-6


def test_jac_structure_positions_match_the_references_python_sparsity(oracle, scipy_model_golden):
    """The set of (row, column) positions of jac_structure! against `bundle_adjustment_sparsity` of the reference's
    Python (src/SolverScipy.py:75-88), converted to the Julia layout; the ORDER of the entries is Julia's own and is
    covered by the structure-formula tests above."""
    g = scipy_model_golden
    npnts, nobs = int(g["shape"][1]), int(g["shape"][2])
    rows, cols = oracle.jac_structure(g["cam_idx"], g["pnt_idx"], npnts)
    assert len(rows) == 24 * nobs
    got = sorted(zip(rows.tolist(), cols.tolist()))
    ref = [tuple(int(v) for v in rc) for rc in g["jac_pattern_rows_cols_1based"].tolist()]
    assert got == ref


def test_jac_coord_thread_chunks_same_values(oracle, ba):
    p = small_problem(ba)
    a = oracle.jac_coord(p.cam_idx, p.pnt_idx, p.x0, p.npnts, nthreads=1)
    b = oracle.jac_coord(p.cam_idx, p.pnt_idx, p.x0, p.npnts, nthreads=3)
    assert np.array_equal(a, b)


# ---- quirks: theta == 0 and z == 0 --------------------------------------
def test_theta_zero_gives_nan_residual_and_zero_block(oracle, ba):
    p = small_problem(ba)
    x = p.x0.copy()
    c = p.cam_idx[5] - 1
    x[3 * p.npnts + 9 * c: 3 * p.npnts + 9 * c + 3] = 0.0
    cx = oracle.cons(p.cam_idx, p.pnt_idx, p.pt2d, x, p.npnts)
    vals = oracle.jac_coord(p.cam_idx, p.pnt_idx, x, p.npnts).reshape(-1, 24)
    hit = p.cam_idx == c + 1
    assert np.all(np.isnan(cx.reshape(-1, 2)[hit]))
    assert np.all(vals[hit] == 0.0)
    assert np.all(np.isfinite(cx.reshape(-1, 2)[~hit]))


def test_z_zero_block_is_zero(oracle):
    # one point, one camera with r along z so that P1.z == X.z + t.z == 0 exactly
    cam, pnt = np.array([1]), np.array([1])
    x = np.array([0.25, -0.5, 1.0, 0.0, 0.0, 0.5, 0.1, 0.2, -1.0, 1e-3, 1e-5, 500.0])
    cx = oracle.cons(cam, pnt, np.zeros(2), x, 1)
    vals = oracle.jac_coord(cam, pnt, x, 1)
    assert not np.all(np.isfinite(cx))
    assert np.all(vals == 0.0)


# ---- mul_sparse (src/lma_aux.jl:194-212; reference test runtests.jl:91-108) -------------------------
def test_mul_sparse_vs_dense():
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    m, n, nnz = 12, 5, 30
    rows = rng.integers(1, m + 1, nnz)
    cols = rng.integers(1, n + 1, nnz)
    vals = rng.normal(size=nnz)
    A = np.zeros((m, n))
    np.add.at(A, (rows - 1, cols - 1), vals)
    x = rng.normal(size=n)
    assert np.allclose(O.mul_sparse(rows, cols, vals, x, m), A @ x, rtol=1e-13, atol=1e-13)
    y = rng.normal(size=m)
    assert np.allclose(O.mul_sparse(cols, rows, vals, y, n), A.T @ y, rtol=1e-13, atol=1e-13)


# ---- LDL of the augmented SQD system (src/ldl_aux.jl) vs dense solve --------------------------------
def test_ldl_solves_sqd_system(oracle):
    import scipy.sparse as sp
    rng = np.random.default_rng(5)
    m, n, lam = 14, 6, 0.7
    J = rng.normal(size=(m, n)) * (rng.random((m, n)) < 0.5)
    K = np.block([[np.eye(m), J], [J.T, -lam * np.eye(n)]])
    U = sp.csc_matrix(np.triu(K))
    U.sort_indices()
    b = rng.normal(size=m + n)
    for P in (None, rng.permutation(m + n)):
        xs = oracle.ldl_solve_csc(m + n, U.indptr, U.indices, U.data, b, P)
        assert np.allclose(K @ xs, b, rtol=1e-11, atol=1e-11)


def test_lm_step_solves_damped_normal_equations(oracle, ba):
    p = small_problem(ba)
    lam = 30.0
    delta, dr2, jtr = oracle.lm_step(p.cam_idx, p.pnt_idx, p.pt2d, p.ncams, p.npnts, p.x0, lam, want_jtr=True)
    rows, cols = oracle.jac_structure(p.cam_idx, p.pnt_idx, p.npnts)
    vals = oracle.jac_coord(p.cam_idx, p.pnt_idx, p.x0, p.npnts)
    r = oracle.cons(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.npnts)
    J = np.zeros((2 * p.nobs, p.nvar))
    np.add.at(J, (rows - 1, cols - 1), vals)
    assert np.allclose(jtr, J.T @ r, rtol=1e-12, atol=1e-9)
    ref = np.linalg.solve(J.T @ J + lam * np.eye(p.nvar), -J.T @ r)
    assert np.linalg.norm(delta - ref) <= 1e-8 * np.linalg.norm(ref)
    assert abs(dr2 - 0.5 * np.linalg.norm(J @ delta + r) ** 2) <= 1e-9 * dr2


def test_lm_loop_reduces_objective_and_follows_lambda_rules(oracle, ba):
    p = small_problem(ba)
    res = oracle.lm_solve(p.cam_idx, p.pnt_idx, p.pt2d, p.ncams, p.npnts, p.x0)
    assert res.status in ("small_step", "first_order", "small_residual", "acceptable")
    f0 = res.log[0]["f"]
    assert res.objective < 0.05 * f0
    # first lambda = max(30, 1e10/||J'r||) (src/lm.jl:59)
    assert res.log[0]["lam"] == max(30.0, 1e10 / res.log[0]["dfeas"])
    for a, b in zip(res.log[:-1], res.log[1:]):
        if a["accepted"]:
            lam = a["lam"] / 3
            if a["rho"] >= 0.9:
                lam /= 3
            assert b["lam"] == max(1e-8, lam)
            assert b["f"] <= a["f"]
        else:
            assert b["lam"] == max(a["lam"], 1 / a["delta_norm"]) * 3
            assert b["f"] == a["f"]


def test_lm_linesearch_variant_runs(oracle, ba):
    p = small_problem(ba)
    res = oracle.lm_solve(p.cam_idx, p.pnt_idx, p.pt2d, p.ncams, p.npnts, p.x0,
                          oracle.default_params(linesearch=1))
    assert res.status != "exception"
    assert res.objective < 0.05 * res.log[0]["f"]


# ---- column normalisation (src/lma_aux.jl:98-178; reference test runtests.jl:31-88) -----------------
def test_normalize_variants_solve_the_same_system(oracle, ba):
    # normalize = :J / :A only equilibrate columns (J D^-1, -lambda D^-2, delta = y / d), so the
    # LM step is the one of (J'J + lambda I) delta = -J'r for every `normalize` argument.  That is why the
    # device path can accept the argument and solve one system.
    p = small_problem(ba)
    lam = 30.0
    rows, cols = oracle.jac_structure(p.cam_idx, p.pnt_idx, p.npnts)
    vals = oracle.jac_coord(p.cam_idx, p.pnt_idx, p.x0, p.npnts)
    r = oracle.cons(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.npnts)
    m, n = 2 * p.nobs, p.nvar
    J = np.zeros((m, n))
    np.add.at(J, (rows - 1, cols - 1), vals)
    ref = np.linalg.solve(J.T @ J + lam * np.eye(n), -J.T @ r)
    d = np.linalg.norm(J, axis=0)                       # normalize_ldl!: norms of the columns of J
    d[d == 0] = 1.0
    Jn = J / d
    K = np.block([[np.eye(m), Jn], [Jn.T, -np.diag(lam / d ** 2)]])   # [[I J D^-1]; [. -lambda D^-2]]
    xr = np.linalg.solve(K, np.concatenate([-r, np.zeros(n)]))
    delta = xr[m:] / d                                   # denormalize_vect!
    assert np.linalg.norm(delta - ref) <= 1e-10 * np.linalg.norm(ref)
    dr = xr[:m]                                          # = -(r + J delta): the model residual of the LDL path
    assert np.linalg.norm(dr + r + J @ ref) <= 1e-9 * np.linalg.norm(r)


def test_ldl_oracle_agrees_with_an_independent_sparse_solver(oracle, ba):
    # the restated ldl_aux.jl against SuperLU on the same augmented SQD system [[I J];[J' -lambda I]]
    # (plays the role of the reference's `A \\ b` checks, test/runtests.jl:111-128, at a realistic size)
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    p = ba.synth.make_problem((16, 600, 3000))
    lam = 75.0
    rows, cols = oracle.jac_structure(p.cam_idx, p.pnt_idx, p.npnts)
    vals = oracle.jac_coord(p.cam_idx, p.pnt_idx, p.x0, p.npnts)
    r = oracle.cons(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.npnts)
    m, n = 2 * p.nobs, p.nvar
    J = sp.csr_matrix((vals, (rows - 1, cols - 1)), shape=(m, n))
    K = sp.bmat([[sp.identity(m), J], [J.T, -lam * sp.identity(n)]], format="csc")
    xr = spla.splu(K).solve(np.concatenate([-r, np.zeros(n)]))
    delta, dr2 = oracle.lm_step(p.cam_idx, p.pnt_idx, p.pt2d, p.ncams, p.npnts, p.x0, lam)
    assert np.linalg.norm(delta - xr[m:]) <= 1e-9 * np.linalg.norm(xr[m:])
    assert abs(dr2 - 0.5 * float(xr[:m] @ xr[:m])) <= 1e-10 * dr2


@pytest.mark.parametrize("shape,lam", [((7, 60, 260), 30.0), ((9, 300, 1500), 1e3), ((12, 400, 2000), 0.5)])
def test_schur_ordered_solve_equals_natural_order_ldl(oracle, shape, lam):
    """The Schur-ordered exact solve (points eliminated first, dense Cholesky of the camera system: the pivot
    order a fill-reducing permutation gives src/ldl_aux.jl on a BA Jacobian) is the same linear solve as the
    natural-order LDL' restated from src/ldl_aux.jl: steps agree to rounding.  It is what the GPU tests use as
    the checker at the BASELINE.json sizes, where the natural order fills in too much."""
    import bundleadjustment.jl_b200.synth as synth
    p = synth.make_problem(shape)
    d0, dr0, j0 = oracle.lm_step(p.cam_idx, p.pnt_idx, p.pt2d, p.ncams, p.npnts, p.x0, lam, want_jtr=True)
    for dense in ("c", "scipy"):
        d1, dr1, j1 = oracle.lm_step_schur(p.cam_idx, p.pnt_idx, p.pt2d, p.ncams, p.npnts, p.x0, lam, want_jtr=True,
                                           dense=dense)
        assert np.array_equal(j0, j1)
        assert np.linalg.norm(d1 - d0) <= 1e-10 * np.linalg.norm(d0)
        assert abs(dr1 - dr0) <= 1e-12 * dr0
    # any observation order (the lists per point are built by a counting sort)
    perm = np.random.default_rng(1).permutation(p.nobs)
    d2, dr2 = oracle.lm_step_schur(p.cam_idx[perm], p.pnt_idx[perm], p.pt2d.reshape(-1, 2)[perm].ravel(), p.ncams,
                                   p.npnts, p.x0, lam, dense="c")
    assert np.linalg.norm(d2 - d0) <= 1e-10 * np.linalg.norm(d0)


def test_lm_loop_with_schur_solver_follows_the_ldl_trajectory(oracle):
    import bundleadjustment.jl_b200.synth as synth
    p = synth.make_problem((9, 300, 1500))
    a = oracle.lm_solve(p.cam_idx, p.pnt_idx, p.pt2d, p.ncams, p.npnts, p.x0)
    b = oracle.lm_solve(p.cam_idx, p.pnt_idx, p.pt2d, p.ncams, p.npnts, p.x0,
                        oracle.default_params(nthreads=2), solver="schur")
    assert a.status == b.status and a.iter == b.iter
    assert [r["accepted"] for r in a.log] == [r["accepted"] for r in b.log]
    for ra, rb in zip(a.log, b.log):
        assert abs(ra["f"] - rb["f"]) <= 1e-9 * ra["f"] and abs(ra["lam"] - rb["lam"]) <= 1e-9 * ra["lam"]
    assert abs(a.objective - b.objective) <= 1e-9 * a.objective


def test_oracle_dense_cholesky(oracle):
    rng = np.random.default_rng(3)
    n = 150
    A = rng.normal(size=(n, n))
    S = A @ A.T + n * np.eye(n)
    b = rng.normal(size=n)
    x = b.copy()
    assert oracle.lib().bao_chol_solve(n, S.copy().reshape(-1), x, 3) == 0
    assert np.linalg.norm(S @ x - b) <= 1e-12 * np.linalg.norm(b)
    S[5, 5] = -1.0
    assert oracle.lib().bao_chol_solve(n, S.copy().reshape(-1), b.copy(), 1) == -1
