"""Multi-GPU (one process per GPU, NCCL over NVLink) parity: the observation-sharded LM path must give the
same step and the same trajectory as one GPU.  Skipped unless two CUDA devices are visible."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q, small_max, solver):
    sys.path.insert(0, ROOT)
    if small_max is not None:
        os.environ["BAGPU_VEC_SMALL_MAX"] = small_max  # read once by libbagpu: selects the PCG vector path
    import torch
    import torch.distributed as dist
    import bundleadjustment.jl_b200 as ba
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = ba.synth.make_problem((12, 400, 2000))
    m = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs, device=rank, rank=rank, nranks=world)
    m.set_solver(solver)
    ba.init_comm(m)
    d, dr2, obj, _, it = ba.lm_step(m, p.x0, 100.0)
    st = ba.Levenberg_Marquardt(m, "LDL", "AMD", "None", False)
    # J'v needs the allreduce as well
    o0, o1 = m.obs_range
    w = np.random.default_rng(1).normal(size=2 * p.nobs)
    jtw = m.jtprod_(p.x0, w[2 * o0:2 * o1])
    if rank == 0:
        q.put(dict(d=d, dr2=dr2, obj=obj, it=it, f=[r["f"] for r in st.rows], acc=[r["accepted"] for r in st.rows],
                   objective=st.objective, status=st.status, x=st.solution, jtw=jtw))
    dist.barrier()
    m.close()
    dist.destroy_process_group()


# PCG with the fused single-CTA vector kernel / with the multi-CTA kernels; the exact solve (sharded assembly of the
# reduced camera system, integer allreduce, replicated factorisation)
@pytest.mark.parametrize("solver,small_max", [("pcg", None), ("pcg", "0"), ("exact", None), ("mixed", None)])
def test_two_gpu_lm_matches_one_gpu(ba, small_max, solver):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1000)
    port += (7 if small_max else 0) + (3 if solver == "exact" else 0) + (5 if solver == "mixed" else 0)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, small_max, solver)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = q.get(timeout=600)
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    p = ba.synth.make_problem((12, 400, 2000))
    m = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs)
    m.set_solver(solver)
    d, dr2, obj, _, it = ba.lm_step(m, p.x0, 100.0)
    st = ba.Levenberg_Marquardt(m, "LDL", "AMD", "None", False)
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    assert rel(got["d"], d) <= 1e-10
    assert abs(got["dr2"] - dr2) <= 1e-11 * dr2 and abs(got["obj"] - obj) <= 1e-12 * obj
    assert got["status"] == st.status and got["acc"] == [r["accepted"] for r in st.rows]
    assert np.allclose(got["f"], [r["f"] for r in st.rows], rtol=1e-9, atol=0)
    assert abs(got["objective"] - st.objective) <= 1e-9 * st.objective
    assert rel(got["x"], st.solution) <= 1e-7
    w = np.random.default_rng(1).normal(size=2 * p.nobs)
    assert rel(got["jtw"], m.jtprod_(p.x0, w)) <= 1e-11


BIG = (160, 10000, 50000)   # 1440 camera rows: multi-CTA vector kernels, PCG long enough to harvest Ritz vectors


def _deflated_sequence(ba, m, p):
    out = []
    for lam in (30.0, 30.0, 3.0):          # harvest; deflated; stale vectors + refreshed ones
        d, dr2, _, _, it = ba.lm_step(m, p.x0, lam, pcg_max_iter=2000)
        out.append((d, dr2, it))
    st = ba.Levenberg_Marquardt(m, "LDL", "AMD", "None", False, ite_max=4, pcg_max_iter=2000)
    return out, st


def _worker_deflated(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["BAGPU_DEFL_CHECK"] = "1"   # the library cross-checks its 4-vector Schur product (coarse setup)
    import torch
    import torch.distributed as dist
    import bundleadjustment.jl_b200 as ba
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = ba.synth.make_problem(BIG)
    m = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs, device=rank, rank=rank, nranks=world)
    m.set_solver("pcg")
    m.set_deflation(32)
    ba.init_comm(m)
    seq, st = _deflated_sequence(ba, m, p)
    if rank == 0:
        q.put(dict(seq=seq, f=[r["f"] for r in st.rows], acc=[r["accepted"] for r in st.rows],
                   pcg=[r["pcg_iters"] for r in st.rows], objective=st.objective))
    dist.barrier()
    m.close()
    dist.destroy_process_group()


def test_two_gpu_deflated_pcg_matches_one_gpu(ba):
    """The deflation vectors are built from replicated data with the same arithmetic on every rank: the sharded
    run must follow the single-GPU one (same steps to PCG accuracy, same iteration counts up to rounding)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1000) + 13
    procs = [ctx.Process(target=_worker_deflated, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = q.get(timeout=600)
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    p = ba.synth.make_problem(BIG)
    m = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs)
    m.set_solver("pcg")
    m.set_deflation(32)
    seq, st = _deflated_sequence(ba, m, p)
    m.set_deflation(0)
    d_plain, _, _, _, it_plain = ba.lm_step(m, p.x0, 30.0, pcg_max_iter=2000)
    m.close()
    rel = lambda a, b: float(np.linalg.norm(a - b) / np.linalg.norm(b))
    print("pcg iterations: plain", it_plain, "1 GPU", [s[2] for s in seq], "2 GPUs", [s[2] for s in got["seq"]],
          "LM", [r["pcg_iters"] for r in st.rows], got["pcg"])
    for (d2, dr2, it2), (d1, dr1, it1) in zip(got["seq"], seq):
        assert rel(d2, d1) <= 1e-9
        assert abs(dr2 - dr1) <= 1e-9 * abs(dr1)
        assert abs(it2 - it1) <= max(3, 0.2 * it1), (it1, it2)
    assert rel(got["seq"][0][0], d_plain) <= 1e-9
    assert 2 * got["seq"][1][2] <= it_plain, (it_plain, got["seq"][1][2])      # the vectors are in use on 2 GPUs
    assert got["acc"] == [r["accepted"] for r in st.rows]
    assert np.allclose(got["f"], [r["f"] for r in st.rows], rtol=1e-8, atol=0)
    assert abs(got["objective"] - st.objective) <= 1e-8 * st.objective


@pytest.mark.parametrize("solver", ["pcg", "exact", "mixed"])
def test_single_process_multi_gpu_handle_matches_one_gpu(ba, oracle, solver):
    """ba_create_multi: several GPUs behind one handle in one process (the mode the reference's single Julia
    process can reach).  Every host-pointer entry point must return the full-length arrays of the one-GPU handle."""
    import torch
    from conftest import TOL, parity_report, rel_errors
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs two CUDA devices")
    ng = min(ng, 4)
    p = ba.synth.make_problem((160, 10000, 50000))   # camera system above the single-CTA threshold
    m1 = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs)
    mg = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs, ngpus=ng)
    assert mg.nobs_local == p.nobs and mg.obs_range == (0, p.nobs)
    for m in (m1, mg):
        m.set_solver(solver)
    cx1, v1 = m1.cons_jac_coord_(p.x0)
    cxg, vg = mg.cons_jac_coord_(p.x0)
    assert np.array_equal(cx1, cxg) and np.array_equal(v1, vg)          # same kernels on the same observations
    assert np.array_equal(m1.cons(p.x0), mg.cons(p.x0))
    r1, c1 = m1.jac_structure()
    rg, cg = mg.jac_structure()
    assert np.array_equal(r1, rg) and np.array_equal(c1, cg)
    rng = np.random.default_rng(2)
    v, w = rng.normal(size=p.nvar), rng.normal(size=2 * p.nobs)
    assert np.array_equal(m1.jprod_(p.x0, v), mg.jprod_(p.x0, v))
    e = rel_errors(mg.jtprod_(p.x0, w), m1.jtprod_(p.x0, w))
    assert e[0] <= 1e-13
    d1, dr1, o1, j1, _ = ba.lm_step(m1, p.x0, 30.0, want_jtr=True)
    dg, drg, og, jg, _ = ba.lm_step(mg, p.x0, 30.0, want_jtr=True)
    es, ej = rel_errors(dg, d1), rel_errors(jg, j1)
    assert es[0] <= TOL and ej[0] <= 1e-13 and abs(drg - dr1) <= 1e-11 * dr1 and abs(og - o1) <= 1e-13 * o1
    s1 = ba.Levenberg_Marquardt(m1, "LDL", "AMD", "None", False, ite_max=5, solver=solver)
    sg = ba.Levenberg_Marquardt(ba.FeasibilityResidual(mg), "LDL", "AMD", "None", False, ite_max=5, solver=solver)
    assert sg.status == s1.status and sg.iter == s1.iter
    assert [r["accepted"] for r in sg.rows] == [r["accepted"] for r in s1.rows]
    worst = max(abs(a["f"] - b["f"]) / b["f"] for a, b in zip(sg.rows, s1.rows))
    parity_report("single_process_multi_gpu", ngpus=ng, solver=solver, step=es[0], jtr=ej[0], trajectory_f=worst,
                  objective=abs(sg.objective - s1.objective) / s1.objective)
    assert worst <= 1e-9 and abs(sg.objective - s1.objective) <= 1e-9 * s1.objective
    assert rel_errors(sg.solution, s1.solution)[0] <= 1e-8
    with pytest.raises(ba.BAError):
        ba._lib.check(ba._lib.lib().ba_set_profiling(mg.handle, 1), mg.handle)   # per-GPU call: refused on a group
    mg.close()
    m1.close()
