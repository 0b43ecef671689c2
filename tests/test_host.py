"""CPU tests of the host side: the C-ABI library loads and exports every symbol include/bagpu.h
declares (no compute calls without a GPU), host-only entry points, and the Python mirror's
argument handling."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, small_problem


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "bagpu.h")).read()
    return sorted(set(re.findall(r"^BA_API [\w\s\*]+?\b(ba_\w+)\(", text, re.M)))


def test_library_exports_every_declared_symbol(ba):
    L = ba._lib.lib()
    names = _declared_symbols()
    assert len(names) >= 28
    for n in names:
        assert getattr(L, n) is not None, n
    assert sorted(ba._lib.SYMBOLS) == names  # the Python binding covers the whole header


def test_version_and_default_params(ba):
    assert b"sm_100a" in ba._lib.lib().ba_version()
    p = ba.default_params()
    eps = np.finfo(np.float64).eps
    assert p.restol == p.ortol == p.rtol == np.cbrt(eps)          # src/lm.jl:21-24
    assert p.satol == p.srtol == p.oatol == p.atol == np.sqrt(eps)
    assert (p.nu_d, p.nu_m, p.lam, p.delta_d, p.ite_max) == (3.0, 3.0, 30.0, 2.0, 200)  # :25-26


def test_partition_observations_cuts_on_point_boundaries(ba):
    p = small_problem(ba)
    L = ba._lib.lib()
    for nr in (1, 2, 3, 8):
        cuts = np.empty(nr + 1, dtype=np.int64)
        rc = L.ba_partition_observations(p.nobs, p.pnt_idx.ctypes.data_as(C.c_void_p), nr,
                                         cuts.ctypes.data_as(C.c_void_p))
        assert rc == 0
        assert cuts[0] == 0 and cuts[-1] == p.nobs and np.all(np.diff(cuts) >= 0)
        for c in cuts[1:-1]:
            assert c == p.nobs or p.pnt_idx[c] != p.pnt_idx[c - 1]
        assert np.diff(cuts).max() <= p.nobs / nr + p.ncams  # balanced up to one point's track
    # not point-major -> BA_ERR_UNSORTED
    bad = p.pnt_idx[::-1].copy()
    cuts = np.empty(3, dtype=np.int64)
    assert L.ba_partition_observations(p.nobs, bad.ctypes.data_as(C.c_void_p), 2,
                                       cuts.ctypes.data_as(C.c_void_p)) == ba._lib.BA_ERR_UNSORTED


def test_name_mangling_like_reference(ba):
    # src/BALNLPModels.jl:58-68
    assert ba.name("LadyBug/problem-49-7776-pre.txt.bz2") == "LadyBug-49-7776"
    assert ba.name("Venice/problem-1778-993923-pre.txt.bz2") == "Venice-1778-993923"


def test_create_rejects_bad_indices_without_gpu(ba):
    # argument validation happens before any CUDA call
    L = ba._lib.lib()
    cam = np.array([1, 3], dtype=np.int64)  # 3 > ncams
    pnt = np.array([1, 1], dtype=np.int64)
    pt = np.zeros(4)
    h = C.c_void_p()
    rc = L.ba_create(2, 1, 2, cam.ctypes.data_as(C.c_void_p), pnt.ctypes.data_as(C.c_void_p),
                     pt.ctypes.data_as(C.c_void_p), 0, C.byref(h))
    assert rc == ba._lib.BA_ERR_ARG
    assert b"out of range" in L.ba_last_error(h)
    L.ba_destroy(h)


def test_synth_problem_shape_and_order(ba):
    p = small_problem(ba)
    assert p.cam_idx.min() >= 1 and p.cam_idx.max() == p.ncams and len(np.unique(p.cam_idx)) == p.ncams
    assert np.all(np.diff(p.pnt_idx) >= 0)                      # point-major
    same = np.diff(p.pnt_idx) == 0
    assert np.all(np.diff(p.cam_idx)[same] > 0)                 # cameras ascending and distinct in a point
    assert np.bincount(p.pnt_idx)[1:].min() >= 2
    assert p.x0.size == 9 * p.ncams + 3 * p.npnts and p.pt2d.size == 2 * p.nobs
    q = small_problem(ba)
    assert np.array_equal(p.pt2d, q.pt2d)                       # deterministic


def test_named_shapes_match_baseline_configs(ba):
    S = ba.synth.SHAPES
    assert S["ladybug-49"] == (49, 7776, 31843)
    assert S["trafalgar-257"] == (257, 65132, 225911)
    assert S["dubrovnik-356"] == (356, 226730, 1255268)
    assert S["venice-1778"] == (1778, 993923, 5001946)
    assert S["final-13682"] == (13682, 4456117, 28987644)


def test_struct_layouts_match_the_header(ba, tmp_path):
    """The ctypes mirrors (and the isbits structs of julia/BALGPUModels.jl, which list the same fields in the
    same order) must have the layout a C compiler gives include/bagpu.h."""
    import subprocess
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "bagpu.h"
#define P(T, f) printf(#T "." #f " %zu\n", offsetof(T, f))
int main(void) {
  printf("ba_lm_params %zu\n", sizeof(ba_lm_params));
  printf("ba_lm_row %zu\n", sizeof(ba_lm_row));
  printf("ba_lm_stats %zu\n", sizeof(ba_lm_stats));
  P(ba_lm_params, restol); P(ba_lm_params, nu_d); P(ba_lm_params, lambda); P(ba_lm_params, ite_max);
  P(ba_lm_params, linesearch); P(ba_lm_params, pcg_max_iter); P(ba_lm_params, pcg_tol); P(ba_lm_params, solver);
  P(ba_lm_row, iter); P(ba_lm_row, rho); P(ba_lm_row, accepted); P(ba_lm_row, ntimes); P(ba_lm_row, solver);
  P(ba_lm_row, converged); P(ba_lm_row, solve_rel);
  P(ba_lm_stats, status); P(ba_lm_stats, iter); P(ba_lm_stats, objective); P(ba_lm_stats, pcg_iters_total);
  P(ba_lm_stats, t_backsub_ms); P(ba_lm_stats, capped_solves); P(ba_lm_stats, worst_solve_rel);
  P(ba_lm_stats, t_prepare_ms); P(ba_lm_stats, t_chol_ms); P(ba_lm_stats, chol_count);
  P(ba_lm_stats, mixed_fallbacks);
  return 0;
}
''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = dict(line.rsplit(" ", 1) for line in subprocess.check_output([str(exe)], text=True).strip().splitlines())
    L = ba._lib
    assert int(got["ba_lm_params"]) == C.sizeof(L.LMParams)
    assert int(got["ba_lm_row"]) == C.sizeof(L.LMRow)
    assert int(got["ba_lm_stats"]) == C.sizeof(L.LMStats)
    for cname, cls, names in (("ba_lm_params", L.LMParams, {"restol": "restol", "nu_d": "nu_d", "lambda": "lam",
                                                             "ite_max": "ite_max", "linesearch": "linesearch",
                                                             "pcg_max_iter": "pcg_max_iter", "pcg_tol": "pcg_tol",
                                                             "solver": "solver"}),
                              ("ba_lm_row", L.LMRow, {"iter": "iter", "rho": "rho", "accepted": "accepted",
                                                       "ntimes": "ntimes", "solver": "solver", "converged": "converged",
                                                       "solve_rel": "solve_rel"}),
                              ("ba_lm_stats", L.LMStats, {"status": "status", "iter": "iter", "objective": "objective",
                                                           "pcg_iters_total": "pcg_iters_total",
                                                           "t_backsub_ms": "t_backsub_ms",
                                                           "capped_solves": "capped_solves",
                                                           "worst_solve_rel": "worst_solve_rel",
                                                           "t_prepare_ms": "t_prepare_ms", "t_chol_ms": "t_chol_ms",
                                                           "chol_count": "chol_count",
                                                           "mixed_fallbacks": "mixed_fallbacks"})):
        for cf, pf in names.items():
            assert int(got["%s.%s" % (cname, cf)]) == getattr(cls, pf).offset, (cname, cf)


def test_plain_c_consumer(ba, tmp_path):
    import subprocess
    lib = ba._lib.LIB_PATH
    exe = tmp_path / "abi_driver"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "abi_driver.c"), "-o", str(exe), lib,
                           "-Wl,-rpath," + os.path.dirname(lib)])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "abi ok" in out.stdout


def test_lanczos_tridiagonal_eigensolver_matches_numpy(ba):
    # CG coefficients of a real PCG run define T; its eigenpairs are the Ritz pairs of the deflation space
    rng = np.random.default_rng(0)
    n = 60
    A = rng.normal(size=(n, n))
    A = A @ A.T + n * np.eye(n)
    b = rng.normal(size=n)
    x, r = np.zeros(n), b.copy()
    p, rz = r.copy(), r @ r
    al, be = [], []
    for _ in range(25):
        q = A @ p
        a = rz / (p @ q)
        x += a * p
        r -= a * q
        rzn = r @ r
        al.append(a)
        be.append(rzn / rz)
        p = r + be[-1] * p
        rz = rzn
    m = len(al)
    al, be = np.array(al), np.array(be)
    T = np.zeros((m, m))
    for j in range(m):
        T[j, j] = 1 / al[j] + (be[j - 1] / al[j - 1] if j else 0.0)
        if j + 1 < m:
            T[j, j + 1] = T[j + 1, j] = -np.sqrt(be[j]) / al[j]
    w_ref, _ = np.linalg.eigh(T)
    w = np.empty(m)
    V = np.empty(m * m)
    rc = ba._lib.lib().ba_dbg_tridiag_eig(al.ctypes.data_as(C.c_void_p), be.ctypes.data_as(C.c_void_p), m,
                                          w.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p))
    assert rc == 0
    assert np.allclose(w, w_ref, rtol=1e-12, atol=1e-12 * abs(w_ref).max())
    V = V.reshape(m, m).T                       # column-major -> V[:, j] eigenvector j
    assert np.allclose(T @ V, V * w, atol=1e-10 * abs(w_ref).max())
    assert np.allclose(V.T @ V, np.eye(m), atol=1e-12)
    # Ritz values of A lie inside its spectrum
    ev = np.linalg.eigvalsh(A)
    assert w.min() >= ev.min() * (1 - 1e-10) and w.max() <= ev.max() * (1 + 1e-10)


def test_select_orthonormal_drops_dependent_columns(ba):
    rng = np.random.default_rng(1)
    Y = rng.normal(size=(200, 6))
    Y = np.concatenate([Y[:, :3], Y[:, :1] * 2.0 + 1e-9 * rng.normal(size=(200, 1)), Y[:, 3:]], axis=1)  # col 3 ~ ghost of col 0
    n = Y.shape[1]
    G = np.ascontiguousarray(Y.T @ Y)
    Cc = np.zeros(n * n)
    kept = C.c_int32()
    rc = ba._lib.lib().ba_dbg_select_columns(G.ctypes.data_as(C.c_void_p), n, 5, 1e-3, Cc.ctypes.data_as(C.c_void_p),
                                             C.byref(kept))
    assert rc == 0 and kept.value == 5
    Cm = Cc[: n * kept.value].reshape(n, kept.value)
    assert np.all(Cm[3] == 0.0)                 # the ghost column takes no part
    Q = Y @ Cm
    assert np.allclose(Q.T @ Q, np.eye(kept.value), atol=1e-10)


def test_julia_glue_structs_list_the_header_fields_in_order():
    """julia/lm_gpu.jl cannot be executed here (no Julia): at least its isbits mirrors of ba_lm_params / ba_lm_row /
    ba_lm_stats must name the fields of include/bagpu.h in the same order with matching widths, and every symbol the
    glue ccalls must be declared in the header."""
    import re
    hdr = open(os.path.join(ROOT, "include", "bagpu.h")).read()
    jl = open(os.path.join(ROOT, "bundleadjustment.jl_b200", "julia", "lm_gpu.jl")).read()
    jl2 = open(os.path.join(ROOT, "bundleadjustment.jl_b200", "julia", "BALNLPModels.jl")).read()

    def c_fields(name):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        out = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            ty, names = decl.split(None, 1)
            for n in names.split(","):
                out.append((n.strip(), {"double": "Float64", "int64_t": "Int64", "int32_t": "Int32"}[ty]))
        return out

    def jl_fields(name):
        body = re.search(r"struct %s\n(.*?)\nend" % name, jl, re.S).group(1)
        return [(m.group(1), m.group(2)) for m in re.finditer(r"(\w+)::(\w+)", body)]

    for cname, jname in (("ba_lm_params", "BALMParams"), ("ba_lm_stats", "BALMStats"), ("ba_lm_row", "BALMRow")):
        assert c_fields(cname) == jl_fields(jname), cname
    declared = set(re.findall(r"BA_API\s+[\w\s\*]+?\b(ba_\w+)\s*\(", hdr))
    used = set(re.findall(r"ccall\(\(:(ba_\w+), libbagpu\)", jl + jl2))
    assert used and used <= declared, used - declared
