"""Multi-GPU (one process per GPU, NCCL over NVLink) parity: the observation-sharded LM path must give the
same step and the same trajectory as one GPU.  Skipped unless two CUDA devices are visible."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q, small_max):
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


@pytest.mark.parametrize("small_max", [None, "0"])  # fused single-CTA vector kernel / multi-CTA kernels
def test_two_gpu_lm_matches_one_gpu(ba, small_max):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port + (7 if small_max else 0), q, small_max)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = q.get(timeout=600)
    for pr in procs:
        pr.join(timeout=120)
        assert pr.exitcode == 0
    p = ba.synth.make_problem((12, 400, 2000))
    m = ba.BALNLPModel(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.ncams, p.npnts, p.nobs)
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
