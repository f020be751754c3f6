"""Multi-GPU plumbing: one process per GPU, torch.distributed only for the rendezvous (it
carries the 128-byte NCCL unique id from rank 0); the gradient all-reduce itself is issued by
libsparkfm_b200.so on its own stream (sfm_comm_init / DESIGN.md section 3.5)."""
from __future__ import annotations

import numpy as np


def shard_range(n_rows: int, rank: int, world: int):
    """Contiguous block sharding of the data set rows: rank r owns
    [n*r/world, n*(r+1)/world)."""
    return n_rows * rank // world, n_rows * (rank + 1) // world


def broadcast_bytes(payload: bytes, n: int, src: int = 0, device="cpu") -> bytes:
    """Broadcast n bytes from rank `src` over the default torch.distributed group."""
    import torch
    import torch.distributed as dist
    if dist.get_rank() == src:
        t = torch.tensor(list(payload), dtype=torch.uint8, device=device)
    else:
        t = torch.zeros(n, dtype=torch.uint8, device=device)
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().astype(np.uint8).tolist())


def init_comm(handle, device="cuda"):
    """Joins `handle` to an NCCL communicator spanning the torch.distributed world."""
    import torch.distributed as dist
    from .handle import Handle
    from ._lib import SFM_UNIQUE_ID_BYTES
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = Handle.comm_unique_id() if rank == 0 else b""
    uid = broadcast_bytes(uid, SFM_UNIQUE_ID_BYTES, 0, device)
    handle.comm_init(uid, rank, world)
    return rank, world
