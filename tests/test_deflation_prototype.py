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
