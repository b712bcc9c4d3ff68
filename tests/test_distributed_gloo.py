"""world_size-2 `gloo` test of the N>1 path on CPU: the product's partitioning rule
(qinchworm_b200.mpi.rank_sub_range = src/mpi.jl:49-54) gives every rank a disjoint Sobol index
range; rank-local partial integrals (produced here by the oracle, standing in for the device) are
combined with one all-reduce per step (src/mpi.jl:104-127) and must equal the single-rank result.
On the GPU box the same reduction is one ncclAllReduce inside libqinchworm_cuda.so."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, N, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import models
    from oracle import oracle as orc
    from qinchworm_b200 import mpi
    ex, grid, f = models.anderson(n_tau=16)
    pl = ex.flatten()
    o = orc.Oracle(pl, ex.P)
    ids = []
    for order in range(0, 3):
        for k in ([0] if order == 0 else range(1, 2 * order)):
            pr, pa = orc.topologies(order, k)
            o.set_topologies(len(ids), orc.MODE_BOLD, order, k, pr, pa)
            ids.append(len(ids))
    rng = mpi.rank_sub_range(N)                  # product rule, world taken from torch.distributed
    assert mpi.world() == (rank, world)
    part = o.eval(0.0, grid.tau[5], grid.tau[6], ids, N, start=rng.start - 1, count=len(rng))
    if rank != 0:
        part[0] = 0                              # order-0 entry is exact and identical on every rank
    t = torch.from_numpy(part.view(np.float64).copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)     # ONE collective for all entries of the step
    total = t.numpy().view(np.complex128)
    if rank == 0:
        full = o.eval(0.0, grid.tau[5], grid.tau[6], ids, N)
        q.put((float(np.abs(total - full).max()), float(np.abs(full).max()), (rng.start, len(rng))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("N", [64, 37])
def test_two_rank_partial_sums_gloo(N):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, N, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, scale, rng0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-13 * max(scale, 1.0)
    assert rng0 == (1, (N + 1) // 2)
