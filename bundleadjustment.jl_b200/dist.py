"""One-process-per-GPU plumbing: torch.distributed carries the NCCL unique id of libbagpu's own
communicator (ba_comm_unique_id / ba_comm_init); the data path never goes through torch."""
from __future__ import annotations

import ctypes as C

from . import _lib


def init_comm(model, group=None, p2p=True):
    """Create the NCCL communicator of a sharded BALNLPModel.  Needs an initialised torch.distributed
    process group (any backend) whose ranks match model.rank / model.nranks.  With ``p2p`` (default) the ranks
    also exchange CUDA IPC handles so that the per-PCG-iteration sum runs over peer memory (NVLink) fused into
    the kernel that consumes it; BAGPU_NO_P2P=1 keeps that exchange on NCCL."""
    import torch
    import torch.distributed as dist
    if model.nranks == 1:
        return
    if not dist.is_initialized():
        raise RuntimeError("init_comm needs torch.distributed to be initialised")
    if dist.get_world_size(group) != model.nranks or dist.get_rank(group) != model.rank:
        raise ValueError("process group does not match the model's rank/nranks")
    L = _lib.lib()
    buf = (C.c_uint8 * 128)()
    if model.rank == 0:
        _lib.check(L.ba_comm_unique_id(buf))
    t = torch.tensor(list(buf), dtype=torch.uint8)
    if dist.get_backend(group) == "nccl":
        t = t.cuda(model.device)
    dist.broadcast(t, 0, group=group)
    arr = (C.c_uint8 * 128)(*t.cpu().tolist())
    _lib.check(L.ba_comm_init(model.handle, arr), model.handle)
    if p2p:
        # peer-memory mailboxes for the per-PCG-iteration exchange: swap CUDA IPC handles (64 bytes per rank).
        # Every rank must end up on the same path, so success is agreed on collectively.
        ok = 1
        try:
            mine = (C.c_uint8 * 64)()
            _lib.check(L.ba_comm_ipc_export(model.handle, mine), model.handle)
        except _lib.BAError:
            ok, mine = 0, (C.c_uint8 * 64)()
        on_gpu = dist.get_backend(group) == "nccl"
        t = torch.tensor(list(mine), dtype=torch.uint8)
        t = t.cuda(model.device) if on_gpu else t
        parts = [torch.empty_like(t) for _ in range(model.nranks)]
        dist.all_gather(parts, t, group=group)
        if ok:
            flat = [b for q in parts for b in q.cpu().tolist()]
            allh = (C.c_uint8 * (64 * model.nranks))(*flat)
            ok = 1 if L.ba_comm_ipc_import(model.handle, allh) == 0 else 0
        flag = torch.tensor([ok], dtype=torch.int32)
        flag = flag.cuda(model.device) if on_gpu else flag
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag.item()) == 0:
            _lib.check(L.ba_comm_ipc_disable(model.handle), model.handle)  # NCCL for everybody
