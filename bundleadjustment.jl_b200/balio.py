"""BAL problem files (https://grail.cs.washington.edu/projects/bal/): reader with the semantics of the
reference's ``readfile`` (src/ReadFiles.jl:9-53) and a writer, so that a synthetic problem can be handed
to the unmodified Julia reference and a real ``problem-*-pre.txt.bz2`` can be handed to this library.

Grammar (one token group per line):
    ncams npnts nobs
    nobs  x  "cam pnt x y"          0-based indices in the file, 1-based in memory (ReadFiles.jl:21-27)
    ncams x  9 lines                 r1 r2 r3 t1 t2 t3 f k1 k2 in the FILE;
                                     (r, t, k1, k2, f) in x0 -- f moves from 7th to 9th (ReadFiles.jl:32-43)
    npnts x  3 lines                 X Y Z
``x0 = [points; cameras]`` (ReadFiles.jl:29-47).  Numbers are written with ``repr`` (shortest string that
round-trips), so write -> read reproduces every FP64 bit.
"""
from __future__ import annotations

import bz2
import io

import numpy as np


def _open(path, mode):
    if str(path).endswith(".bz2"):
        return io.TextIOWrapper(bz2.open(path, mode + "b"), encoding="ascii")
    return open(path, mode)


def readfile(filename):
    """-> (cam_indices, pnt_indices, pt2d, x0, ncams, npnts, nobs), exactly what src/ReadFiles.jl:9 returns
    (indices 1-based int64, pt2d interleaved, x0 = [points; cameras] with camera = (r, t, k1, k2, f))."""
    with _open(filename, "r") as f:
        ncams, npnts, nobs = (int(t) for t in f.readline().split())
        obs = np.loadtxt(f, max_rows=nobs, dtype=np.float64, ndmin=2) if nobs else np.zeros((0, 4))
        rest = np.loadtxt(f, dtype=np.float64, ndmin=1) if (ncams or npnts) else np.zeros(0)
    if obs.shape != (nobs, 4):
        raise ValueError("expected %d observation lines 'cam pnt x y'" % nobs)
    if rest.size != 9 * ncams + 3 * npnts:
        raise ValueError("expected %d parameter lines, found %d" % (9 * ncams + 3 * npnts, rest.size))
    cam_indices = obs[:, 0].astype(np.int64) + 1
    pnt_indices = obs[:, 1].astype(np.int64) + 1
    pt2d = np.ascontiguousarray(obs[:, 2:4]).ravel()
    cams_file = rest[: 9 * ncams].reshape(ncams, 9)           # r t f k1 k2
    cams = np.concatenate([cams_file[:, 0:6], cams_file[:, 7:9], cams_file[:, 6:7]], axis=1)  # r t k1 k2 f
    x0 = np.concatenate([rest[9 * ncams:], cams.ravel()])
    return cam_indices, pnt_indices, pt2d, x0, ncams, npnts, nobs


def writefile(filename, cam_indices, pnt_indices, pt2d, x, ncams, npnts):
    """Inverse of ``readfile``: 1-based indices in, 0-based on disk; camera (r,t,k1,k2,f) -> file order
    (r,t,f,k1,k2).  ``.bz2`` names are compressed like the BAL distribution."""
    cam_indices = np.asarray(cam_indices, dtype=np.int64)
    pnt_indices = np.asarray(pnt_indices, dtype=np.int64)
    pt2d = np.asarray(pt2d, dtype=np.float64).reshape(-1, 2)
    x = np.asarray(x, dtype=np.float64)
    nobs = cam_indices.size
    if x.size != 9 * ncams + 3 * npnts or pt2d.shape[0] != nobs or pnt_indices.size != nobs:
        raise ValueError("inconsistent sizes")
    cams = x[3 * npnts:].reshape(ncams, 9)
    cams_file = np.concatenate([cams[:, 0:6], cams[:, 8:9], cams[:, 6:8]], axis=1)
    with _open(filename, "w") as f:
        f.write("%d %d %d\n" % (ncams, npnts, nobs))
        lines = ["%d %d %s %s\n" % (c - 1, p - 1, repr(float(u)), repr(float(v)))
                 for c, p, (u, v) in zip(cam_indices.tolist(), pnt_indices.tolist(), pt2d.tolist())]
        f.write("".join(lines))
        f.write("".join(repr(float(v)) + "\n" for v in cams_file.ravel().tolist()))
        f.write("".join(repr(float(v)) + "\n" for v in x[: 3 * npnts].tolist()))


def write_problem(filename, p):
    """Write a synth.BALProblem (its starting point x0)."""
    writefile(filename, p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts)
