"""CPU test (gloo, world_size 2) of the multi-GPU decomposition: observations cut on point boundaries by
ba_partition_observations; per-observation outputs are slices; point blocks are owned by exactly one rank;
camera blocks and the camera part of J'r are sums over ranks (the buffers libbagpu allreduces over NCCL).
The arithmetic here is the oracle's (this is a test of the decomposition, not of the kernels)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import bundleadjustment.jl_b200 as ba
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    p = ba.synth.make_problem((7, 60, 260))
    cuts = np.empty(world + 1, dtype=np.int64)
    rc = ba._lib.lib().ba_partition_observations(p.nobs, p.pnt_idx.ctypes.data_as(C.c_void_p), world,
                                                 cuts.ctypes.data_as(C.c_void_p))
    assert rc == 0
    o0, o1 = int(cuts[rank]), int(cuts[rank + 1])
    cam, pnt, pt = p.cam_idx[o0:o1], p.pnt_idx[o0:o1], p.pt2d[2 * o0:2 * o1]
    # local pieces
    r_loc = O.cons(cam, pnt, pt, p.x0, p.npnts)
    vals = O.jac_coord(cam, pnt, p.x0, p.npnts).reshape(-1, 2, 12)
    A, B = vals[:, :, :3], vals[:, :, 3:]
    U = np.zeros((p.ncams, 9, 9))
    np.add.at(U, cam - 1, np.einsum("kia,kib->kab", B, B))
    gc = np.zeros((p.ncams, 9))
    np.add.at(gc, cam - 1, -np.einsum("kia,ki->ka", B, r_loc.reshape(-1, 2)))
    V = np.zeros((p.npnts, 3, 3))
    np.add.at(V, pnt - 1, np.einsum("kia,kib->kab", A, A))
    f2 = np.array([float(r_loc @ r_loc)])
    tU, tg, tV, tf = (torch.from_numpy(a.copy()) for a in (U, gc, V, f2))
    owned = torch.from_numpy((np.abs(V).sum(axis=(1, 2)) > 0).astype(np.int64))
    for t in (tU, tg, tV, tf, owned):
        dist.all_reduce(t)  # sum
    if rank == 0:
        # full problem on one rank
        r = O.cons(p.cam_idx, p.pnt_idx, p.pt2d, p.x0, p.npnts)
        v = O.jac_coord(p.cam_idx, p.pnt_idx, p.x0, p.npnts).reshape(-1, 2, 12)
        A, B = v[:, :, :3], v[:, :, 3:]
        U0 = np.zeros((p.ncams, 9, 9))
        np.add.at(U0, p.cam_idx - 1, np.einsum("kia,kib->kab", B, B))
        g0 = np.zeros((p.ncams, 9))
        np.add.at(g0, p.cam_idx - 1, -np.einsum("kia,ki->ka", B, r.reshape(-1, 2)))
        V0 = np.zeros((p.npnts, 3, 3))
        np.add.at(V0, p.pnt_idx - 1, np.einsum("kia,kib->kab", A, A))
        ok = (np.allclose(tU.numpy(), U0, rtol=1e-12, atol=1e-9) and np.allclose(tg.numpy(), g0, rtol=1e-12, atol=1e-9)
              and np.allclose(tV.numpy(), V0, rtol=1e-12, atol=1e-12) and abs(tf.item() - float(r @ r)) <= 1e-10 * float(r @ r)
              and int(owned.max()) == 1)  # every point block lives on exactly one rank: no point-side exchange
        q.put(bool(ok))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_decomposition_sums_to_the_full_problem():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    ok = q.get(timeout=240)
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    assert ok
