"""Multi-GPU plumbing: one process per GPU (torch.distributed for rendezvous, barriers and reductions), one pmk context per
process.  The data path between GPUs is the library's own NCCL all-gather of each wavefront step's mutations; this module only
ships the 128-byte NCCL unique id from rank 0 to the other ranks and hands it to pmk_comm_init."""
from __future__ import annotations

import numpy as np

from . import pmk


def broadcast_bytes(dist, payload: bytes | None, nbytes: int, src: int = 0, device=None) -> bytes:
    """Broadcast a fixed-size byte string with whatever backend `dist` was initialised with (gloo on CPU, nccl on GPUs)."""
    import torch
    t = torch.zeros(nbytes, dtype=torch.uint8, device=device or "cpu")
    if dist.get_rank() == src:
        t.copy_(torch.from_numpy(np.frombuffer(payload, np.uint8).copy()))
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


def init_comm(ctx: "pmk.Context", dist=None, device=None):
    """rank/size from torch.distributed (or single process); creates the library's communicator over the same ranks."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        ctx.comm_init(0, 1, None)
        return 0, 1
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = pmk.comm_unique_id() if rank == 0 else None
    uid = broadcast_bytes(dist, uid, 128, 0, device)
    ctx.comm_init(rank, world, uid)
    return rank, world


def step_tasks(gw: int, gh: int, diag: int, rank: int, nranks: int):
    """dest cells (x, y) of anti-diagonal `diag` of one view that rank `rank` sweeps: the cells of the diagonal are numbered from its
    smallest x, rank r takes the numbers G with G % nranks == r -- the same arithmetic as the sweep driver (one view per step)."""
    xlo, xhi = max(0, diag - gh + 1), min(gw - 1, diag)
    cells = [(x, diag - x) for x in range(xlo, xhi + 1) if (x - xlo) % nranks == rank]
    assert len(cells) == pmk.step_share(max(0, xhi - xlo + 1), rank, nranks)
    return cells
