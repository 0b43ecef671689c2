"""CPU prototype (numpy, dense reduced camera system) of the next solver step planned in DESIGN.md section 9:
deflating the block-Jacobi PCG with Ritz vectors HARVESTED FROM A PCG SOLVE ITSELF (CG is Lanczos: no extra
products with S), and re-using them after lambda and x have changed, as consecutive LM iterations do.

This is a model of a planned device feature, kept as a test so that the claim it rests on stays checked:
the spectrum of blkdiag(S)^-1 S has ~1 % small outliers (gauge-like, smooth modes) and is tightly clustered
around 1 otherwise; removing 32 harvested Ritz vectors cuts the iterations >= 3x, also with stale vectors.
The arithmetic (Jacobian blocks) comes from the oracle; nothing here touches the product path."""
import numpy as np
import pytest
import scipy.sparse as sp


def _blocks(oracle, p, x):
    vals = oracle.jac_coord(p.cam_idx, p.pnt_idx, x, p.npnts, nthreads=4).reshape(-1, 2, 12)
    F = oracle.cons(p.cam_idx, p.pnt_idx, p.pt2d, x, p.npnts, nthreads=4).reshape(-1, 2)
    pi, ci = p.pnt_idx - 1, p.cam_idx - 1
    rows = np.repeat(np.arange(2 * p.nobs), 12)
    cols = np.concatenate([(3 * pi)[:, None] + np.arange(3), (9 * ci)[:, None] + np.arange(9) + 3 * p.npnts], axis=1)
    cols = np.repeat(cols[:, None, :], 2, axis=1).reshape(-1)
    J = sp.csr_matrix((vals.reshape(-1), (rows, cols)), shape=(2 * p.nobs, p.nvar))
    Jp, Jc = J[:, :3 * p.npnts].tocsc(), J[:, 3 * p.npnts:].tocsc()
    Vb = np.zeros((p.npnts, 3, 3))
    np.add.at(Vb, pi, np.einsum("kia,kib->kab", vals[:, :, :3], vals[:, :, :3]))
    g = -(J.T @ F.reshape(-1))
    return (Jc.T @ Jc).toarray(), (Jc.T @ Jp).tocsr(), Vb, g[:3 * p.npnts], g[3 * p.npnts:]


def _schur(U, W, Vb, gp, gc, lam, npnts, ncams):
    inv = np.linalg.inv(Vb + lam * np.eye(3))
    i = (3 * np.arange(npnts))[:, None, None] + np.arange(3)[None, :, None] + np.zeros((1, 1, 3), int)
    j = (3 * np.arange(npnts))[:, None, None] + np.arange(3)[None, None, :] + np.zeros((1, 3, 1), int)
    Vinv = sp.csr_matrix((inv.reshape(-1), (i.reshape(-1), j.reshape(-1))), shape=(3 * npnts, 3 * npnts))
    S = U + lam * np.eye(9 * ncams) - (W @ Vinv @ W.T).toarray()
    Mb = np.stack([np.linalg.inv(S[9 * c:9 * c + 9, 9 * c:9 * c + 9]) for c in range(ncams)])
    return S, gc - W @ (Vinv @ gp), Mb


def _bj(Mb, r):
    return np.einsum("cab,cb->ca", Mb, r.reshape(-1, 9)).reshape(-1)


def _pcg(S, b, prec, tol=1e-13, maxit=3000, harvest=False):
    x, r = np.zeros_like(b), b.copy()
    z = prec(r)
    p, rz = z.copy(), r @ z
    rz0, Z, al, be = rz, [], [], []
    for it in range(1, maxit + 1):
        if harvest:
            Z.append(z / np.sqrt(rz))          # M-orthonormal Lanczos basis (up to the sign (-1)^j)
        q = S @ p
        a = rz / (p @ q)
        x += a * p
        r -= a * q
        z = prec(r)
        rzn = r @ z
        al.append(a)
        if np.sqrt(rzn / rz0) <= tol:
            break
        be.append(rzn / rz)
        p = z + be[-1] * p
        rz = rzn
    return it, (np.array(Z).T if harvest else None), np.array(al), np.array(be)


def _ritz(Z, al, be, k):
    """k Ritz vectors of M^-1 S from the CG coefficients: T = Z' S Z is tridiagonal with
    T_jj = 1/a_j + b_{j-1}/a_{j-1},  T_{j,j+1} = -sqrt(b_j)/a_j  (minus: Z is not sign-alternated)."""
    m = Z.shape[1]
    T = np.zeros((m, m))
    for j in range(m):
        T[j, j] = 1 / al[j] + (be[j - 1] / al[j - 1] if j else 0.0)
        if j + 1 < m:
            T[j, j + 1] = T[j + 1, j] = -np.sqrt(be[j]) / al[j]
    _, s = np.linalg.eigh(T)
    Y = Z @ s[:, :4 * k]                       # candidates; converged Ritz values come with ghost copies
    keep = []
    for j in range(Y.shape[1]):                # Gram-Schmidt drops the ghosts
        v = Y[:, j].copy()
        for u in keep:
            v -= (u @ v) * u
        if np.linalg.norm(v) > 1e-3 * np.linalg.norm(Y[:, j]):
            keep.append(v / np.linalg.norm(v))
        if len(keep) == k:
            break
    return np.array(keep).T


@pytest.mark.timeout(300)
def test_harvested_ritz_vectors_deflate_pcg_and_survive_an_lm_step(oracle, ba):
    p = ba.synth.make_problem((120, 8000, 40000))
    U, W, Vb, gp, gc = _blocks(oracle, p, p.x0)
    lam = 100.0
    S, b, Mb = _schur(U, W, Vb, gp, gc, lam, p.npnts, p.ncams)
    it0, Z, al, be = _pcg(S, b, lambda r: _bj(Mb, r), harvest=True)
    Y = _ritz(Z, al, be, 32)
    Ac = np.linalg.inv(Y.T @ S @ Y)
    it1 = _pcg(S, b, lambda r: _bj(Mb, r) + Y @ (Ac @ (Y.T @ r)))[0]
    assert it1 * 3 <= it0, (it0, it1)
    # the next LM iteration: x moved, lambda divided by 9 -- same (now stale) vectors, new coarse matrix
    x1 = p.x0 + 0.3 * (p.x_true - p.x0)
    U1, W1, Vb1, gp1, gc1 = _blocks(oracle, p, x1)
    S1, b1, Mb1 = _schur(U1, W1, Vb1, gp1, gc1, lam / 9, p.npnts, p.ncams)
    it2 = _pcg(S1, b1, lambda r: _bj(Mb1, r))[0]
    Ac1 = np.linalg.inv(Y.T @ S1 @ Y)
    it3 = _pcg(S1, b1, lambda r: _bj(Mb1, r) + Y @ (Ac1 @ (Y.T @ r)))[0]
    assert it3 * 3 <= it2, (it2, it3)


def _c_ritz(ba, Z, al, be, ncand, k, base=None, full_below=160):
    """The device feature's host half, through the library's own helpers: smallest Ritz pairs of the CG
    tridiagonal (ba_dbg_tridiag_smallest), candidates Y = [base, Z s], Gram matrix, ghost-free selection
    (ba_dbg_select_columns), new basis = Y C."""
    import ctypes as C
    L = ba._lib.lib()
    m = min(Z.shape[1], len(be))
    ncand = min(ncand, m)
    al, be = np.ascontiguousarray(al[:m]), np.ascontiguousarray(be[:m])
    w, V = np.empty(ncand), np.empty(m * ncand)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert L.ba_dbg_tridiag_smallest(vp(al), vp(be), m, ncand, full_below, vp(w), vp(V)) == 0
    Y = Z[:, :m] @ V.reshape(ncand, m).T
    if base is not None:
        Y = np.concatenate([base, Y], axis=1)
    n = Y.shape[1]
    G = np.ascontiguousarray(Y.T @ Y)
    Cc, kept = np.zeros(n * n), C.c_int32()
    assert L.ba_dbg_select_columns(vp(G), n, k, 1e-3, vp(Cc), C.byref(kept)) == 0
    return Y @ Cc[: n * kept.value].reshape(n, kept.value), w


@pytest.mark.timeout(600)
def test_library_ritz_helpers_reproduce_the_prototype(oracle, ba):
    """Same experiment with the C++ helpers the device path uses (both eigen-solver branches), on top of the
    cluster coarse space, plus the rolling refresh: 32 base vectors from the first solve, 16 more from each
    deflated solve, while lambda falls by 9x per step."""
    p = ba.synth.make_problem((120, 8000, 40000))
    n9, ncl = 9 * p.ncams, 16
    cpc = -(-p.ncams // ncl)
    P = np.zeros((n9, 6 * ncl))
    for c in range(p.ncams):
        P[9 * c + np.arange(6), (c // cpc) * 6 + np.arange(6)] = 1
    P = P[:, P.any(0)]

    def two(S, Mb, B):
        Ai = np.linalg.inv(B.T @ S @ B)
        return lambda r: _bj(Mb, r) + B @ (Ai @ (B.T @ r))

    lam, base, roll, counts = 100.0, None, None, []
    for t in (0.0, 0.3, 0.6, 0.8):
        x = p.x0 + t * (p.x_true - p.x0)
        S, b, Mb = _schur(*_blocks(oracle, p, x), lam, p.npnts, p.ncams)
        if base is None:
            it0, Z, al, be = _pcg(S, b, two(S, Mb, P), harvest=True)
            base, w = _c_ritz(ba, Z, al, be, 64, 32)
            base_ii, w_ii = _c_ritz(ba, Z, al, be, 64, 32, full_below=0)   # inverse-iteration branch
            assert base.shape[1] == 32 and base_ii.shape[1] == 32
            assert np.allclose(w, w_ii, rtol=1e-9)
            # same subspace from both branches
            assert np.linalg.norm(base_ii - base @ (base.T @ base_ii)) < 1e-6
            assert np.allclose(base.T @ base, np.eye(32), atol=1e-9)
        B = np.concatenate([P, base] + ([roll] if roll is not None else []), axis=1)
        it, Z2, al2, be2 = _pcg(S, b, two(S, Mb, B), harvest=True)
        it_plain = _pcg(S, b, two(S, Mb, P))[0]
        counts.append((it_plain, it))
        full, _ = _c_ritz(ba, Z2, al2, be2, 32, 48, base=base)
        assert np.allclose(full[:, :32], base, atol=1e-9)                  # the base is kept as it is
        roll = full[:, 32:]
        lam /= 9
    for it_plain, it in counts:
        assert it * 2.5 <= it_plain, counts


def test_inverse_iteration_handles_ghost_ritz_values(ba):
    """CG run far past convergence: the tridiagonal has ghost copies of converged Ritz values; the vectors
    returned for them must still be eigenvectors (residual) even if nearly parallel."""
    import ctypes as C
    rng = np.random.default_rng(3)
    n = 80
    Q, _ = np.linalg.qr(rng.normal(size=(n, n)))
    ev = np.concatenate([[1e-3, 2e-3, 5e-3], np.linspace(0.5, 1.5, n - 3)])
    A = (Q * ev) @ Q.T
    b = rng.normal(size=n)
    x, r = np.zeros(n), b.copy()
    p, rz = r.copy(), r @ r
    al, be = [], []
    for _ in range(240):                          # 3n steps: orthogonality is long lost
        q = A @ p
        a = rz / (p @ q)
        x += a * p
        r -= a * q
        rzn = r @ r
        if rzn < 1e-290:
            break
        al.append(a)
        be.append(rzn / rz)
        p = r + be[-1] * p
        rz = rzn
    m, k = len(al), 24
    al, be = np.array(al), np.array(be)
    T = np.zeros((m, m))
    for j in range(m):
        T[j, j] = 1 / al[j] + (be[j - 1] / al[j - 1] if j else 0.0)
        if j + 1 < m:
            T[j, j + 1] = T[j + 1, j] = -np.sqrt(be[j]) / al[j]
    w, V = np.empty(k), np.empty(m * k)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert ba._lib.lib().ba_dbg_tridiag_smallest(vp(al), vp(be), m, k, 0, vp(w), vp(V)) == 0
    V = V.reshape(k, m).T
    wr = np.linalg.eigvalsh(T)[:k]
    assert np.allclose(w, wr, rtol=1e-8, atol=1e-12 * abs(T).max())
    res = np.linalg.norm(T @ V - V * w, axis=0)
    assert np.all(res < 1e-8 * abs(T).max()), res
    assert np.allclose(np.linalg.norm(V, axis=0), 1.0)
