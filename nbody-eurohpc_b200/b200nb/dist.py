"""torch.distributed plumbing for the one-process-per-GPU launch (torchrun): rendezvous, the ncclUniqueId broadcast
and the slice arithmetic.  Plumbing only — the position exchange itself is the library's own ncclAllGather
(csrc/context.cu: enqueue_gather), the analogue of the reference's MPI_Allgatherv
(src/murb/implem/SimulationNBodyMultiNode.cpp:93-117)."""
from __future__ import annotations

import os

from . import slice_length


def env_rank() -> tuple[int, int, int]:
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend: str | None = None):
    """Initialise torch.distributed from the torchrun environment (MASTER_ADDR defaults to 127.0.0.1)."""
    import torch
    import torch.distributed as dist

    rank, world, local_rank = env_rank()
    if world == 1:
        return None
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if backend is None:
        backend = "cpu:gloo,cuda:nccl" if torch.cuda.is_available() else "gloo"
    if torch.cuda.is_available():
        torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return dist


def broadcast_bytes(dist, payload: bytes | None, src: int = 0) -> bytes:
    """Rank `src` passes the payload (the 128-byte ncclUniqueId), every rank returns it."""
    box = [payload]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def slice_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Global body range [first, last) owned by `rank`: contiguous, L = slice_length(n, world) targets per rank
    (buildCountsDispls analogue, SimulationNBodyMultiNode.cpp:76-91)."""
    L = slice_length(n, world)
    first = min(rank * L, n)
    return first, min(first + L, n)
