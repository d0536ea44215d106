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


def band_partition(gheight: int, nranks: int):
    """[(ylo, yhi)] per rank: contiguous row bands that tile [0, gheight) (pmk_band_rows)."""
    return [pmk.band_rows(gheight, r, nranks) for r in range(nranks)]


def step_tasks(gw: int, gh: int, diag: int, ylo: int, yhi: int):
    """dest cells (x range) of anti-diagonal `diag` whose row lies in [ylo, yhi): the same arithmetic as the sweep driver."""
    xlo = max(max(0, diag - gh + 1), diag - yhi + 1)
    xhi = min(min(gw - 1, diag), diag - ylo)
    return (xlo, xhi) if xhi >= xlo else None
