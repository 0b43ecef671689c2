"""CPU model (numpy / scipy, dense reduced camera system from the oracle) of the mixed-precision solver of DESIGN.md
section 5c (the reference's facto_type < T mode, src/lm.jl:92-98,165-173), kept as a test so
that the claims the device path rests on stay checked without a GPU:

 * an FP32 Cholesky factor of the Jacobi-scaled reduced camera system preconditions FP64 CG on the FP64 operator so
   well that a handful of iterations reach the FP64 solution (the step keeps the 1e-10 bar of north_star);
 * the three-term TF32 split the tensor-core factorisation uses (a = hi + lo, both TF32, rounded by integer
   arithmetic exactly as split_tf32_fast in ba_chol.cu; a b' ~ lo hi' + hi lo' + hi hi') is FP32-accurate;
 * when the preconditioner is useless the iteration cap is reached and the caller can tell (fall-back to FP64).
The arithmetic (Jacobian values, Schur complement) comes from the oracle; nothing here touches the product path."""
import numpy as np
import pytest
import scipy.linalg as sla


def _schur_system(oracle, p, lam, nt=4):
    r = oracle.cons(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.npnts, nt)
    vals = oracle.jac_coord(p.cam_idx, p.pnt_idx, p.x0, p.npnts, nt)
    rows, cols = oracle.jac_structure(p.cam_idx, p.pnt_idx, p.npnts)
    jtr = oracle.mul_sparse(cols, rows, vals, r, p.nvar)
    n9 = 9 * p.ncams
    S, b = np.empty((n9, n9)), np.empty(n9)
    Vinv, hp = np.empty(9 * p.npnts), np.empty(3 * p.npnts)
    oracle.lib().bao_schur_system(p.cam_idx, p.pnt_idx, vals, jtr, p.ncams, p.npnts, p.nobs, float(lam), nt,
                                  S.reshape(-1), b, Vinv, hp)
    return S, b


def _mixed_solve(S, b, tol=1e-13, max_cg=30, quantize=None):
    """The device algorithm (ba_lm.cu mixed_solve): Jacobi scaling, low-precision factor of the scaled matrix, FP64 CG on
    S with M^-1 = D^-1 (L L')^-1 D^-1.  Returns (x, iterations, converged)."""
    d = np.sqrt(np.diag(S))
    St = (S / d[:, None] / d[None, :]).astype(np.float32)
    if quantize is not None:
        St = quantize(St)
    L = sla.cholesky(St, lower=True, check_finite=False).astype(np.float64)

    def M(r):
        y = sla.solve_triangular(L, r / d, lower=True, check_finite=False)
        return sla.solve_triangular(L.T, y, lower=False, check_finite=False) / d

    x = np.zeros_like(b)
    r = b.copy()
    z = M(r)
    p = z.copy()
    rz = float(r @ z)
    bn = float(np.linalg.norm(b))
    prev = rel = 1.0
    for it in range(1, max_cg + 1):
        q = S @ p
        a = rz / float(p @ q)
        x += a * p
        r -= a * q
        prev, rel = rel, float(np.linalg.norm(r)) / bn
        if it >= 2 and (rel <= tol or (rel <= 1e-10 and rel > 0.25 * prev)):
            return x, it, True
        z = M(r)
        rzn = float(r @ z)
        p = z + (rzn / rz) * p
        rz = rzn
    return x, max_cg, False


@pytest.mark.parametrize("shape,lam", [("ladybug-49", 30.0), ("ladybug-49", 0.37), ((60, 3000, 15000), 30.0),
                                        ((60, 3000, 15000), 1e-2)])
def test_fp32_factor_preconditions_fp64_cg_to_the_fp64_solution(oracle, ba, shape, lam):
    p = ba.synth.make_problem(shape)
    S, b = _schur_system(oracle, p, lam)
    x_ref = sla.cho_solve(sla.cho_factor(S, lower=True), b)
    x, it, ok = _mixed_solve(S, b)
    err = float(np.linalg.norm(x - x_ref) / np.linalg.norm(x_ref))
    res = float(np.linalg.norm(b - S @ x) / np.linalg.norm(b))
    assert ok and it <= 12, (it, ok)
    assert err <= 1e-10 and res <= 1e-11, (err, res)


def test_a_useless_preconditioner_hits_the_cap(oracle, ba):
    """A factor of a matrix far too inaccurate (entries rounded to FP16: the scaled matrix loses its definiteness
    margin) must not be mistaken for a converged solve: either the factorisation fails or CG stops at its cap, and
    the caller falls back to FP64."""
    p = ba.synth.make_problem("ladybug-49")
    S, b = _schur_system(oracle, p, 1e-3)
    try:
        x, it, ok = _mixed_solve(S, b, max_cg=8, quantize=lambda A: A.astype(np.float16).astype(np.float32))
    except (np.linalg.LinAlgError, sla.LinAlgError, ValueError):
        return  # non-positive pivot: the device path's first fall-back route
    if ok:  # it did converge: then it must be the right solution
        x_ref = sla.cho_solve(sla.cho_factor(S, lower=True), b)
        assert np.linalg.norm(x - x_ref) <= 1e-9 * np.linalg.norm(x_ref)
    else:
        assert it == 8


def _split_tf32_fast(x):
    """split_tf32_fast of ba_chol.cu: hi = x rounded to TF32 (ties away) by integer arithmetic, lo = the remainder
    rounded the same way; both have their 13 low mantissa bits zero."""
    x = np.asarray(x, dtype=np.float32)
    hi = ((x.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    rem = (x - hi).astype(np.float32)
    lo = ((rem.view(np.uint32) + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
    return hi, lo


def test_three_term_tf32_split_is_fp32_accurate():
    rng = np.random.default_rng(0)
    a = (rng.normal(size=(96, 256)) * np.exp(rng.normal(size=(96, 256)))).astype(np.float32)
    b = (rng.normal(size=(80, 256)) * np.exp(rng.normal(size=(80, 256)))).astype(np.float32)
    ah, al = _split_tf32_fast(a)
    bh, bl = _split_tf32_fast(b)
    for h, l, x in ((ah, al, a), (bh, bl, b)):
        assert not np.any(h.view(np.uint32) & np.uint32(0x1FFF)) and not np.any(l.view(np.uint32) & np.uint32(0x1FFF))
        assert np.max(np.abs((h.astype(np.float64) + l.astype(np.float64)) - x) / np.abs(x)) <= 2.0 ** -21.9
    exact = a.astype(np.float64) @ b.astype(np.float64).T
    three = (al.astype(np.float64) @ bh.astype(np.float64).T + ah.astype(np.float64) @ bl.astype(np.float64).T +
             ah.astype(np.float64) @ bh.astype(np.float64).T)
    one = ah.astype(np.float64) @ bh.astype(np.float64).T
    scale = np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64).T
    e3, e1 = np.max(np.abs(three - exact) / scale), np.max(np.abs(one - exact) / scale)
    assert e3 <= 2.0 ** -20 and e1 >= 50 * e3, (e3, e1)   # three terms: FP32 level; one TF32 term: ~2^-11
