"""Deterministic synthetic BAL-shaped problems (the BAL datasets are not available offline).

Shapes follow the reference's data list (get_data.sh:17-78) and BASELINE.json's configs; the layout
of everything returned is the reference's own (src/ReadFiles.jl:9-53, src/BALNLPModels.jl:79-88):

* ``cam_idx`` / ``pnt_idx``: 1-based int64, one per observation, **point-major** order as in BAL
  files (all observations of point 1 first, cameras ascending within a point; cf. the golden case
  test/runtests.jl:15-17 whose 5 observations all belong to point 1);
* ``pt2d``: interleaved (x, y) per observation;
* ``x0 = [X_1 .. X_npnts (3 each) ; C_1 .. C_ncams (9 each)]``, camera = (r, t, k1, k2, f).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

# name -> (ncams, npnts, nobs); SURVEY.md section 8 table
SHAPES = {
    "ladybug-49": (49, 7776, 31843),
    "trafalgar-257": (257, 65132, 225911),
    "dubrovnik-356": (356, 226730, 1255268),
    "venice-1778": (1778, 993923, 5001946),
    "final-13682": (13682, 4456117, 28987644),
}
_SEEDS = {n: 0xBA10000 + i for i, n in enumerate(SHAPES)}


@dataclass
class BALProblem:
    name: str
    ncams: int
    npnts: int
    nobs: int
    cam_idx: np.ndarray  # int64, 1-based
    pnt_idx: np.ndarray  # int64, 1-based
    pt2d: np.ndarray     # float64, 2*nobs
    x0: np.ndarray       # float64, 3*npnts + 9*ncams
    x_true: np.ndarray

    @property
    def nvar(self):
        return 9 * self.ncams + 3 * self.npnts


def _rodrigues_matrix(r):
    th = np.sqrt((r * r).sum(-1))
    k = r / th[:, None]
    c, s = np.cos(th), np.sin(th)
    K = np.zeros(r.shape[:-1] + (3, 3))
    K[:, 0, 1], K[:, 0, 2] = -k[:, 2], k[:, 1]
    K[:, 1, 0], K[:, 1, 2] = k[:, 2], -k[:, 0]
    K[:, 2, 0], K[:, 2, 1] = -k[:, 1], k[:, 0]
    eye = np.eye(3)[None]
    return c[:, None, None] * eye + s[:, None, None] * K + (1 - c)[:, None, None] * (k[:, :, None] * k[:, None, :])


def project(X, cams):
    """Vectorised BAL projection (same formula as src/BALNLPModels.jl:17-33); X (n,3), cams (n,9)."""
    R = _rodrigues_matrix(cams[:, 0:3])
    P1 = np.einsum("nij,nj->ni", R, X) + cams[:, 3:6]
    P2 = -P1[:, :2] / P1[:, 2:3]
    n2 = (P2 * P2).sum(-1)
    sf = 1.0 + cams[:, 6] * n2 + cams[:, 7] * n2 * n2
    return (cams[:, 8] * sf)[:, None] * P2


def _degrees(rng, npnts, nobs, ncams):
    """Per-point observation counts: 2 + Poisson(mean-2), clipped to [2, dmax], summing to nobs."""
    dmax = min(ncams, 4096)
    mean = nobs / npnts
    d = 2 + rng.poisson(max(mean - 2.0, 0.05), size=npnts)
    # a heavy tail like real tracks: 0.5% of the points are seen by many more cameras
    heavy = rng.random(npnts) < 0.005
    d[heavy] += rng.geometric(1.0 / max(4.0 * mean, 8.0), size=int(heavy.sum()))
    d = np.clip(d, 2, dmax).astype(np.int64)
    diff = int(nobs - d.sum())
    while diff != 0:  # distribute the remainder one observation at a time over random points
        step = 1 if diff > 0 else -1
        cand = np.flatnonzero((d < dmax) if step > 0 else (d > 2))
        pick = rng.choice(cand, size=min(abs(diff), len(cand)), replace=False)
        d[pick] += step
        diff = int(nobs - d.sum())
    return d


def make_problem(shape: str | tuple = "ladybug-49", seed: int | None = None, stress: bool = False,
                 big_rotations: bool = False, noise_px: float = 0.5) -> BALProblem:
    """Generate a BAL-shaped problem.  ``stress`` uses k1,k2 large enough to exercise the distortion
    columns numerically; ``big_rotations`` draws |r| up to ~1.3 rad (golden cameras 4-5 of
    test/runtests.jl:18 have |r| = 1.22)."""
    if isinstance(shape, str):
        name = shape
        ncams, npnts, nobs = SHAPES[shape]
        seed = _SEEDS[shape] if seed is None else seed
    else:
        ncams, npnts, nobs = shape
        name = "custom-%d-%d-%d" % (ncams, npnts, nobs)
        seed = 0xBA1FFFF if seed is None else seed
    assert nobs >= 2 * npnts and ncams >= 2, "every point needs >= 2 observations"
    rng = np.random.Generator(np.random.PCG64(seed))

    # ---- geometry: points in a slab around the centroid (0,0,-2); cameras look at it -------------
    centroid = np.array([0.0, 0.0, -2.0])
    X = np.empty((npnts, 3))
    X[:, 0:2] = rng.uniform(-1.0, 1.0, size=(npnts, 2))
    X[:, 2] = rng.uniform(-2.5, -1.5, size=npnts)
    if big_rotations:
        r = rng.normal(0.0, 0.1, size=(ncams, 3))
        r[:, 1] += rng.uniform(-1.3, 1.3, size=ncams)
    else:
        r = rng.normal(0.0, 0.1, size=(ncams, 3))
    R = _rodrigues_matrix(r)
    depth = rng.uniform(2.5, 4.0, size=ncams)
    t = np.stack([rng.normal(0, 0.05, ncams), rng.normal(0, 0.1, ncams), -depth], axis=1) \
        - np.einsum("nij,j->ni", R, centroid)
    f = rng.uniform(390.0, 1100.0, size=ncams)
    if stress:
        k1 = rng.normal(0.0, 0.05, size=ncams)
        k2 = rng.normal(0.0, 0.01, size=ncams)
    else:
        k1 = rng.normal(0.0, 3e-7, size=ncams)
        k2 = rng.normal(0.0, 6e-13, size=ncams)
    cams = np.concatenate([r, t, k1[:, None], k2[:, None], f[:, None]], axis=1)

    # ---- visibility graph, point-major -----------------------------------------------------------
    deg = _degrees(rng, npnts, nobs, ncams)
    pnt0 = np.repeat(np.arange(npnts, dtype=np.int64), deg)          # 0-based point of each obs
    first = np.concatenate([[0], np.cumsum(deg)[:-1]])
    within = np.arange(nobs, dtype=np.int64) - np.repeat(first, deg)  # 0..deg-1 inside a point
    # window start per point: spatially coherent (x coordinate) with popularity skew on big sets
    pos = (X[:, 0] + 1.0) / 2.0
    if ncams > 1000:
        pos = pos ** 2.0  # Zipf-like popularity: low camera ids are seen by many more points
    start = np.floor(pos * ncams).astype(np.int64) + rng.integers(0, max(ncams // 50, 1), size=npnts)
    # distinct cameras inside the window: cumulative positive gaps, wrapped mod ncams
    maxgap = np.maximum((ncams - 1) // np.maximum(deg, 1), 1)
    gaps = 1 + (rng.random(nobs) * np.minimum(np.repeat(maxgap, deg), 6)).astype(np.int64)
    gaps = np.minimum(gaps, np.repeat(maxgap, deg))
    cg = np.cumsum(gaps)
    offs = cg - np.repeat(cg[first], deg)  # 0 for the first observation of a point, then increasing
    cam0 = (np.repeat(start, deg) + offs) % ncams
    # cameras ascending within a point (BAL order)
    order = np.lexsort((cam0, pnt0))
    cam0 = cam0[order]
    # every camera gets at least one observation: hand unused cameras to random distinct points
    used = np.zeros(ncams, dtype=bool)
    used[cam0] = True
    missing = np.flatnonzero(~used)
    if len(missing):
        pts = rng.choice(npnts, size=len(missing), replace=False)
        cam0[first[pts]] = missing  # replace the first observation of those points
        order2 = np.lexsort((cam0, pnt0))
        cam0 = cam0[order2]

    pred = project(X[pnt0], cams[cam0])
    pt2d = pred + rng.normal(0.0, noise_px, size=pred.shape)

    x_true = np.concatenate([X.ravel(), cams.ravel()])
    X0 = X + rng.normal(0.0, 0.02, size=X.shape)
    cams0 = cams.copy()
    cams0[:, 0:6] += rng.normal(0.0, 1e-3, size=(ncams, 6))
    cams0[:, 8] *= 1.0 + rng.normal(0.0, 0.01, size=ncams)
    x0 = np.concatenate([X0.ravel(), cams0.ravel()])
    return BALProblem(name, ncams, npnts, nobs, (cam0 + 1).astype(np.int64), (pnt0 + 1).astype(np.int64),
                      np.ascontiguousarray(pt2d.ravel()), x0, x_true)
