"""Sample partitioning and rank plumbing — host mirror of src/mpi.jl and src/utility.jl:164-179.

One process per GPU.  `torch.distributed` carries the NCCL unique id to all ranks (the role
MPI.bcast plays for a Julia host); the data-path all-reduce itself runs inside
libqinchworm_cuda.so (one ncclAllReduce per inchworm step over NVLink)."""
from __future__ import annotations

import os

import numpy as np

__all__ = ["split_count", "range_from_chunks_and_idx", "rank_sub_range", "ismaster", "world", "init_comm"]


def split_count(N: int, n: int):
    """Vector of n integers that are approximately equal and sum to N (src/utility.jl:164-167)."""
    q, r = divmod(N, n)
    return [q + 1 if i < r else q for i in range(n)]


def range_from_chunks_and_idx(chunk_sizes, idx: int):
    """1-based inclusive range of the idx-th (1-based) chunk (src/utility.jl:175-179)."""
    sidx = 1 + sum(chunk_sizes[:idx - 1])
    return range(sidx, sidx + chunk_sizes[idx - 1])


def world():
    """(rank, world_size) from torch.distributed if initialised, else from the torchrun env."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def rank_sub_range(N: int, rank=None, size=None):
    """Sub-range of 1:N owned by this rank, as a 1-based inclusive range (src/mpi.jl:49-54)."""
    if rank is None or size is None:
        rank, size = world()
    return range_from_chunks_and_idx(split_count(N, size), rank + 1)


def ismaster():
    return world()[0] == 0


def init_comm(ctx, peer=None):
    """Join a lib.Context to the job: the NCCL communicator (fallback path) and, unless disabled
    (peer=False or QIW_NO_PEER=1), the peer-memory mailboxes that let the step kernel perform the
    all-reduce inside its own tail over NVLink.  Returns "peer", "nccl" or "single"."""
    import torch.distributed as dist
    from . import lib
    rank, size = world()
    if size == 1:
        return "single"
    uid = lib.comm_unique_id() if rank == 0 else np.zeros(lib.UNIQUE_ID_BYTES, dtype=np.uint8)
    box = [uid.tobytes()]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(size, rank, np.frombuffer(box[0], dtype=np.uint8))
    if peer is None:
        peer = os.environ.get("QIW_NO_PEER", "0") != "1"
    if not peer:
        return "nccl"
    mine = ctx.peer_handle().tobytes()
    allh = [None] * size
    dist.all_gather_object(allh, mine)
    ok = 1
    try:
        ctx.peer_init(size, rank, np.frombuffer(b"".join(allh), dtype=np.uint8))
    except lib.QiwError:
        ok = 0
    flags = [None] * size
    dist.all_gather_object(flags, ok)
    if not all(flags):
        # no peer access between some GPUs (or CUDA IPC not permitted): every rank leaves the peer path,
        # the NCCL all-reduce of qiw_comm_init carries the collective
        ctx.peer_disable()
        return "nccl"
    return "peer"
