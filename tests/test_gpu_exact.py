"""GPU tests of the exact damped solve: the device Cholesky on its own against
LAPACK, and the explicit reduced camera system + Cholesky + refinement against the PCG path and the oracle."""
import numpy as np
import pytest

from conftest import TOL, parity_report, rel_errors

pytestmark = pytest.mark.gpu


def _spd(n, seed, cond=1e4):
    rng = np.random.default_rng(seed)
    Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
    w = np.logspace(0, np.log10(cond), n)
    return (Q * w) @ Q.T


@pytest.mark.parametrize("n", [1, 50, 128, 129, 700, 2313])
def test_device_cholesky_matches_lapack(ba, n):
    A = _spd(n, n)
    A = 0.5 * (A + A.T)
    b = np.random.default_rng(1).normal(size=n)
    x, L, f_ms, s_ms = ba.lm.dbg_chol(A, b, want_L=True)
    Lr = np.linalg.cholesky(A)
    eL = np.linalg.norm(np.tril(L) - Lr) / np.linalg.norm(Lr)
    res = np.linalg.norm(A @ x - b) / np.linalg.norm(b)
    ex = np.linalg.norm(x - np.linalg.solve(A, b)) / np.linalg.norm(x)
    parity_report("device_cholesky_vs_lapack", n=n, L=eL, residual=res, x=ex, factor_ms=f_ms, solve_ms=s_ms)
    assert eL <= 1e-12 and res <= 1e-11 and ex <= 1e-10


@pytest.mark.parametrize("n", [1, 50, 129, 700, 2313, 3400])
def test_device_fp32_cholesky_is_an_fp32_accurate_factor(ba, n):
    """The mixed-precision factor (FP32 storage, three-TF32-term tensor-core products; n = 3400 takes the two-panel
    updates): backward error and distance to LAPACK's spotrf at the FP32 level, and a preconditioner application."""
    A = _spd(n, n)
    A = 0.5 * (A + A.T)
    b = np.random.default_rng(1).normal(size=n)
    x, L, f_ms, s_ms = ba.lm.dbg_chol(A, b, want_L=True, fp32=True)
    L = np.tril(L)
    Lr = np.linalg.cholesky(A.astype(np.float32)).astype(np.float64)
    back = np.linalg.norm(L @ L.T - A) / np.linalg.norm(A)
    back_ref = np.linalg.norm(Lr @ Lr.T - A) / np.linalg.norm(A)
    eL = np.linalg.norm(L - Lr) / np.linalg.norm(Lr)
    res = np.linalg.norm(A @ x - b) / np.linalg.norm(b)
    parity_report("device_fp32_cholesky_vs_spotrf", n=n, backward=back, backward_spotrf=back_ref, L=eL, residual=res,
                  factor_ms=f_ms, solve_ms=s_ms)
    assert back <= 2e-6 and eL <= 1e-4 and res <= 5e-3      # cond(A) = 1e4: residual ~ cond * 1e-7


@pytest.mark.parametrize("fp32", [False, True])
def test_persistent_sweep_kernel_matches_the_per_step_kernels(ba, monkeypatch, fp32):
    """Both substitution sweeps as one cooperative kernel (k_chol_sweep) against the per-step kernels that remain the
    route for systems of more than 148 tile rows: same solution up to the order of the FP64 sums."""
    n = 1500
    A = _spd(n, 7)
    A = 0.5 * (A + A.T)
    b = np.random.default_rng(2).normal(size=n)
    x1 = ba.lm.dbg_chol(A, b, fp32=fp32)[0]
    monkeypatch.setenv("BAGPU_SWEEP_STEPS", "1")
    x2 = ba.lm.dbg_chol(A, b, fp32=fp32)[0]
    assert np.linalg.norm(x1 - x2) <= 1e-11 * np.linalg.norm(x2)


def test_device_cholesky_flags_a_non_positive_pivot(ba):
    A = _spd(300, 3)
    A[200, 200] = -1.0
    with pytest.raises(ba.BAError) as e:
        ba.lm.dbg_chol(A, np.ones(300))
    assert e.value.code == ba._lib.BA_ERR_NUMERIC


@pytest.mark.parametrize("shape,lam", [((160, 10000, 50000), 30.0), ((160, 10000, 50000), 1e-3),
                                        ("trafalgar-257", 100.0)])
def test_exact_and_pcg_solve_the_same_system(ba, shape, lam):
    p = ba.synth.make_problem(shape)
    m = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs)
    m.set_solver("pcg")
    d0, dr0, _, _, it0 = ba.lm_step(m, p.x0, lam, pcg_max_iter=5000)
    i0 = ba.lm.last_solve_info(m)
    m.set_solver("exact")
    d1, dr1, _, _, it1 = ba.lm_step(m, p.x0, lam)
    i1 = ba.lm.last_solve_info(m)
    d2, dr2, _, _, _ = ba.lm_step(m, p.x0, lam)
    m.close()
    e = rel_errors(d1, d0)
    parity_report("exact_vs_pcg", shape=str(shape), lam=lam, norm=e[0], floor=e[1], entry=e[2], pcg_iters=int(it0),
                  pcg_rel=i0["rel"], direct_solve_rel=i1["rel"])
    assert i0["solver"] == "pcg" and i1["solver"] == "exact" and i0["converged"] and i1["converged"]
    assert e[0] <= (TOL if lam >= 1 else 1e-7)
    assert abs(dr1 - dr0) <= 1e-10 * dr0
    assert np.array_equal(d1, d2) and dr1 == dr2      # fixed-point assembly + ordered factorisation: bit-identical reruns


@pytest.mark.parametrize("shape,lam", [((160, 10000, 50000), 30.0), ("trafalgar-257", 100.0), ("trafalgar-257", 1.0),
                                        ("dubrovnik-356", 30.0)])
def test_mixed_and_exact_solve_the_same_system(ba, shape, lam):
    """BA_SOLVER_MIXED (FP32 tensor-core factor + FP64 CG on the FP64 operator, src/lm.jl:92-98 facto_type) returns
    the FP64 step: same bar as the exact solver."""
    p = ba.synth.make_problem(shape)
    m = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs)
    m.set_solver("exact")
    d0, dr0, _, _, _ = ba.lm_step(m, p.x0, lam)
    m.set_solver("mixed")
    d1, dr1, _, _, it1 = ba.lm_step(m, p.x0, lam)
    i1 = ba.lm.last_solve_info(m)
    d2, dr2, _, _, _ = ba.lm_step(m, p.x0, lam)
    m.close()
    e = rel_errors(d1, d0)
    parity_report("mixed_vs_exact", shape=str(shape), lam=lam, norm=e[0], floor=e[1], entry=e[2], cg_iters=int(it1),
                  true_residual=i1["rel"])
    assert i1["solver"] == "mixed" and i1["converged"] and 1 <= it1 <= 30
    assert e[0] <= TOL and abs(dr1 - dr0) <= 1e-10 * dr0 and i1["rel"] <= 1e-11
    assert np.array_equal(d1, d2) and dr1 == dr2      # deterministic: fixed-point assembly, ordered sums


def test_mixed_solver_falls_back_to_the_fp64_factorisation(ba, monkeypatch):
    """When CG with the FP32 factor does not converge within its cap the solve is redone with the FP64 factorisation
    (never an inexact step): forced here by a cap of one iteration."""
    monkeypatch.setenv("BAGPU_MIXED_MAX_CG", "1")
    p = ba.synth.make_problem((160, 10000, 50000))
    m = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs)
    monkeypatch.delenv("BAGPU_MIXED_MAX_CG")
    m.set_solver("exact")
    d0, dr0, _, _, _ = ba.lm_step(m, p.x0, 30.0)
    m.set_solver("mixed")
    d1, dr1, _, _, _ = ba.lm_step(m, p.x0, 30.0)
    info = ba.lm.last_solve_info(m)
    assert info["solver"] == "exact" and info["converged"]
    assert np.array_equal(d0, d1) and dr0 == dr1
    st = ba.Levenberg_Marquardt(m, "LDL", "AMD", "None", False, ite_max=2, solver="mixed")
    assert st.mixed_fallbacks == st.iter and all(r["solver"] == "exact" for r in st.rows)
    m.close()


@pytest.mark.parametrize("lam", [1e-4, 1e-8])
def test_mixed_solver_at_tiny_lambda_never_returns_an_inexact_step(ba, lam):
    """Far below the dampings LM uses the FP32 factor degrades (many CG iterations) or fails (non-positive pivot):
    whichever route the solve takes -- CG with the FP32 factor, or the FP64 fall-back -- the step is the exact
    solver's up to the conditioning of the system."""
    p = ba.synth.make_problem((160, 10000, 50000))
    m = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs)
    m.set_solver("exact")
    d0, dr0, _, _, _ = ba.lm_step(m, p.x0, lam)
    m.set_solver("mixed")
    d1, dr1, _, _, it = ba.lm_step(m, p.x0, lam)
    info = ba.lm.last_solve_info(m)
    m.close()
    e = rel_errors(d1, d0)
    parity_report("mixed_vs_exact_tiny_lambda", lam=lam, norm=e[0], solver=info["solver"], iters=int(it), rel=info["rel"])
    assert info["solver"] in ("mixed", "exact") and info["converged"]
    assert e[0] <= 1e-6 and abs(dr1 - dr0) <= 1e-9 * abs(dr0)


def test_exact_solver_reports_indefinite_system_as_exception(ba):
    # theta == 0 on one camera -> NaN blocks -> the factorisation meets a non-positive (NaN) pivot -> status exception,
    # the SQDException route of the reference (src/ldl_aux.jl:199, src/lm.jl:401)
    p = ba.synth.make_problem((9, 300, 1500))
    x0 = p.x0.copy()
    x0[3 * p.npnts: 3 * p.npnts + 3] = 0.0
    m = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs)
    st = ba.Levenberg_Marquardt(m, "LDL", "AMD", "None", False, x=x0, solver="exact")
    assert st.status == "exception"


def test_auto_picks_exact_for_small_camera_systems_and_pcg_for_large(ba):
    p = ba.synth.make_problem((12, 400, 2000))
    m = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs)
    ba.lm_step(m, p.x0, 30.0)
    assert ba.lm.last_solve_info(m)["solver"] == "exact"
    m.close()
    # dense systems from 8192 camera unknowns on: the FP32 tensor-core factor + FP64 CG
    p = ba.synth.make_problem((920, 30000, 150000))
    m = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs)
    ba.lm_step(m, p.x0, 30.0)
    assert ba.lm.last_solve_info(m)["solver"] == "mixed"
    m.close()


def test_pcg_iteration_cap_is_reported(ba):
    """A solve that stops at pcg_max_iter is flagged (ba_last_solve_info, ba_lm_row.converged,
    ba_lm_stats.capped_solves) instead of being returned as if it had converged."""
    p = ba.synth.make_problem((160, 10000, 50000))
    m = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs)
    m.set_solver("pcg")
    ba.lm_step(m, p.x0, 30.0, pcg_max_iter=5)
    info = ba.lm.last_solve_info(m)
    assert info["solver"] == "pcg" and not info["converged"] and info["iters"] == 5 and info["rel"] > 1e-13
    st = ba.Levenberg_Marquardt(m, "LDL", "AMD", "None", False, ite_max=1, pcg_max_iter=5, solver="pcg")
    assert st.capped_solves == st.iter and all(not r["converged"] for r in st.rows)
    assert st.worst_solve_rel > 1e-13
    st = ba.Levenberg_Marquardt(m, "LDL", "AMD", "None", False, ite_max=1, pcg_max_iter=5000, solver="pcg")
    assert st.capped_solves == 0 and all(r["converged"] for r in st.rows) and st.worst_solve_rel <= 1e-13
    m.close()
