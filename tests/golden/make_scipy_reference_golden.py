"""Golden residuals AND Jacobian values from the reference's own Python implementation of the BAL model.

Run in the build container only (needs /root/reference):  python tests/golden/make_scipy_reference_golden.py

Source: /root/reference/src/SolverScipy.py:34-72 -- `rotate`, `project`, `fun`: the residual of the same camera model
as src/BALNLPModels.jl:17-33 (Rodrigues rotation, translation, perspective divide, radial distortion k1/k2, focal
length), written independently of the Julia code (scipy-cookbook lineage).  The function definitions are taken from the
reference file as they are (the file's benchmark script below them, which reads the BAL data sets, is not executed).

What is generated, on a small deterministic BAL-shaped problem (bundleadjustment.jl_b200/synth.py, incl. large rotations):
 * `residuals`: `fun(params)` in float64 -- pins `cons!` beyond the five observations of test/runtests.jl;
 * `jac_pattern_rows_cols_1based`: the sparsity pattern of `bundle_adjustment_sparsity` (src/SolverScipy.py:75-88) in the
   Julia layout -- pins the SET of positions `jac_structure!` produces (their order is Julia's own);
 * `jac_vals`: the 2 x 12 Jacobian block of every observation in the layout of `jac_coord!` (row 1 then row 2; columns
   point x, y, z, then camera r, t, k1, k2, f -- src/BALNLPModels.jl:161-206), obtained from the reference's `fun` by
   central differences evaluated in 80-bit extended precision (numpy longdouble): truncation ~ h^2 and rounding
   ~ 1e-19 / h are both far below 1e-10, which the script verifies by repeating the differences with h / 2.
   This pins the HAND-DERIVED Jacobian (src/JacobianByHand.jl) to reference-held code, to ~1e-11, where the reference's
   own tests hold no vector for it.
Layouts are converted both ways: the reference's Python uses params = [cameras (r, t, f, k1, k2); points] and 0-based
indices (the BAL file order), the Julia model x = [points; cameras (r, t, k1, k2, f)] and 1-based indices
(src/ReadFiles.jl:29-47).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
SRC = "/root/reference/src/SolverScipy.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_scipy_model.json")


def reference_functions():
    text = open(SRC).read()
    head = text.split("\nimport time")[0]          # the function definitions; not the benchmark script below them
    ns = {}
    exec(compile(head, SRC, "exec"), ns)
    return ns["fun"], ns["bundle_adjustment_sparsity"]


def to_reference_layout(p, x, dtype):
    npnts, ncams = p.npnts, p.ncams
    X = np.asarray(x[: 3 * npnts], dtype=dtype).reshape(npnts, 3)
    C = np.asarray(x[3 * npnts:], dtype=dtype).reshape(ncams, 9)
    cams_file_order = np.concatenate([C[:, 0:6], C[:, 8:9], C[:, 6:8]], axis=1)  # (r, t, f, k1, k2)
    return np.concatenate([cams_file_order.reshape(-1), X.reshape(-1)])


def main():
    import bundleadjustment.jl_b200.synth as synth
    fun, sparsity = reference_functions()
    p = synth.make_problem((6, 40, 150), big_rotations=True)
    ci, pi = p.cam_idx - 1, p.pnt_idx - 1
    pts2d = p.pt2d.reshape(-1, 2)

    def f(params, dtype):
        return fun(params, p.ncams, p.npnts, ci, pi, pts2d.astype(dtype))

    res = f(to_reference_layout(p, p.x0, np.float64), np.float64)

    LD = np.longdouble
    base = to_reference_layout(p, p.x0, LD)
    nvar = base.size

    def jac(hrel):
        J = np.empty((2 * p.nobs, nvar), dtype=LD)
        for j in range(nvar):
            h = LD(hrel) * max(LD(1), abs(base[j]))
            a, b = base.copy(), base.copy()
            a[j] += h
            b[j] -= h
            J[:, j] = (f(a, LD) - f(b, LD)) / (2 * h)
        return J

    J1, J2 = jac(1e-7), jac(5e-8)
    scale = np.abs(J1).max(axis=1, keepdims=True)
    agree = float((np.abs(J1 - J2) / scale).max())
    assert agree <= 1e-11, agree
    # per observation: columns of its point (reference layout: after the 9 ncams camera parameters) and of its camera
    # (reference order r, t, f, k1, k2 -> Julia order r, t, k1, k2, f)
    vals = np.empty((p.nobs, 2, 12))
    cam_perm = [0, 1, 2, 3, 4, 5, 7, 8, 6]
    for k in range(p.nobs):
        pc = 9 * p.ncams + 3 * pi[k] + np.arange(3)
        cc = 9 * ci[k] + np.array(cam_perm)
        cols = np.concatenate([pc, cc])
        vals[k, 0] = J1[2 * k, cols].astype(np.float64)
        vals[k, 1] = J1[2 * k + 1, cols].astype(np.float64)
    # the sparsity pattern the reference's Python hands to scipy (src/SolverScipy.py:75-88), converted to the Julia
    # layout (x = [points; cameras], camera columns r, t, k1, k2, f; 1-based rows and columns) as a sorted list of pairs
    A = sparsity(p.ncams, p.npnts, ci, pi).tocoo()
    inv_perm = {ref: jl for jl, ref in enumerate(cam_perm)}   # reference camera slot -> Julia camera slot
    pairs = []
    for r_, c_ in zip(A.row.tolist(), A.col.tolist()):
        if c_ < 9 * p.ncams:
            cam, slot = divmod(c_, 9)
            col = 3 * p.npnts + 9 * cam + inv_perm[slot]
        else:
            col = c_ - 9 * p.ncams
        pairs.append((r_ + 1, col + 1))
    pairs.sort()
    fx = {
        "source": "src/SolverScipy.py:34-72 (rotate, project, fun), imported from /root/reference",
        "shape": [p.ncams, p.npnts, p.nobs], "synth": "make_problem((6, 40, 150), big_rotations=True)",
        "cam_idx": p.cam_idx.tolist(), "pnt_idx": p.pnt_idx.tolist(), "pt2d": p.pt2d.tolist(), "x": p.x0.tolist(),
        "residuals": np.asarray(res, dtype=np.float64).tolist(),
        "jac_vals": vals.reshape(-1).tolist(),
        "jac_pattern_rows_cols_1based": pairs,
        "jac_how": "central differences of the reference's fun() in 80-bit extended precision, h = 1e-7 max(1, |x_j|); "
                   "h and h/2 agree to %.1e relative to the largest entry of a row" % agree,
    }
    with open(OUT, "w") as fo:
        json.dump(fx, fo)
    print("wrote", OUT, "agreement of the two step sizes: %.2e" % agree)


if __name__ == "__main__":
    main()
