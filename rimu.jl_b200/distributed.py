"""Multi-GPU bootstrap: one process per GPU (torchrun), determinant space partitioned by address
hash.  Replaces Rimu's MPI bootstrap (mpi_helpers.jl:9-87): torch.distributed is used ONLY to
broadcast the NCCL unique id; the spawn exchange itself is issued by librimu_b200.so."""
from __future__ import annotations

import ctypes as C
import os

from . import _lib
from .hamiltonians import Context, _contexts


def init_distributed(words: int, records_per_peer: int = 1 << 22, table_slots: int | None = None, fresh: bool = False) -> Context:
    """Create this rank's context for `words`-word addresses and attach an NCCL communicator.
    `fresh=True` makes an additional, independent context (own stream, working memory and communicator) --
    one per replica when several independent vectors are advanced concurrently (n_replicas,
    projector_monte_carlo_problem.jl:152,205)."""
    import torch
    import torch.distributed as dist

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    ctx = None if fresh else _contexts.get(words)
    if ctx is None:
        ctx = Context(words, device=local, table_slots=table_slots)
        if not fresh:
            _contexts[words] = ctx
    if world == 1 or ctx.nranks == world:
        return ctx
    if not dist.is_initialized():
        dist.init_process_group("gloo")
    buf = C.create_string_buffer(128)
    if rank == 0:
        _lib.check(_lib.lib().rimu_comm_unique_id(buf))
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    dist.broadcast(t, src=0)
    ctx.attach_comm(bytes(t.tolist()), rank, world, records_per_peer)
    return ctx
