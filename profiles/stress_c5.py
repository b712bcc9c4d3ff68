"""C5 (BASELINE.json configs[4]): Anderson model, orders 0:6, n_tau = 400, N_samples = 2^20 sharded over the
GPUs of the job (strong scaling: the total number of samples is fixed).  Times `steps` bold inchworm steps in
the middle of the run through qiw_eval (one launch + fused peer all-reduce per step) and prints one JSON line.

    python profiles/stress_c5.py [log2_N] [steps]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 profiles/stress_c5.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist
import models
from qinchworm_b200 import lib, mpi
from qinchworm_b200.inchworm import Solver, _bold_entries

log2N = int(sys.argv[1]) if len(sys.argv) > 1 else 20
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N = 2 ** log2N
ex, grid, f = models.anderson(n_tau=400)
ctx = lib.Context(device=local)
solver = Solver(ex, ctx=ctx)
kind = mpi.init_comm(ctx) if world > 1 else "single"
t0 = time.perf_counter()
bold = _bold_entries(solver, range(0, 7), N, None, None)
t_compile = time.perf_counter() - t0
ids = [t.entry_id for t in bold]
st = [ctx.entry_stats(i) for i in ids]
tops = sum(s["n_top"] for s in st)
flops = sum(s["flops_per_sample"] for s in st)
tau = grid.tau
peak = ctx.measure_fp64_peak()
ctx.eval(0.0, tau[200], tau[201], ids, N)            # warm-up: fills the simplex-root cache
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
dev_ms = []
for k in range(steps):
    ctx.eval(0.0, tau[200 + k], tau[201 + k], ids, N)
    dev_ms.append(ctx.last_device_ms())
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / steps
if world > 1:
    t = torch.tensor([wall, float(np.mean(dev_ms))], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall, dev = float(t[0]), float(t[1])
else:
    dev = float(np.mean(dev_ms))
if rank == 0:
    print(json.dumps({"workload": "C5 stress: Anderson orders 0:6, n_tau=400, N_samples=2^%d, one bold step" % log2N, "n_gpus": world,
                      "collective": kind, "topologies": tops, "configurations": sum(s["n_leaves"] for s in st),
                      "compile_s": t_compile, "step_wall_ms": wall * 1e3, "step_device_ms_max": dev,
                      "diagram_evals_per_s": N * tops / wall, "algorithmic_tflops": flops * N / wall / 1e12,
                      "frac_of_measured_fp64_peak_per_gpu": flops * N / wall / 1e12 / peak / world,
                      "projected_full_run_s": wall * 398}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
ctx.close()
