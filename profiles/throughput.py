"""Throughput of one bold inchworm step (C1 model, orders 0:max_order) as a function of N_samples."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import models
from qinchworm_b200 import lib
from qinchworm_b200.inchworm import Solver, _bold_entries
max_order = int(sys.argv[1]) if len(sys.argv) > 1 else 4
Ns = [int(x) for x in sys.argv[2:]] or [2 ** 10, 2 ** 14, 2 ** 17, 2 ** 20]
ex, grid, f = models.anderson(n_tau=200)
ctx = lib.Context(device=0)
peak = ctx.measure_fp64_peak()
print("measured FP64 FMA peak %.2f TFLOP/s" % peak)
solver = Solver(ex, ctx=ctx)
tau = grid.tau
for N in Ns:
    bold = _bold_entries(solver, range(0, max_order + 1), N, None, None)
    ids = [t.entry_id for t in bold]
    st = [ctx.entry_stats(i) for i in ids]
    flops = sum(s["flops_per_sample"] for s in st)
    tops = sum(s["n_top"] for s in st)
    leaves = sum(s["n_leaves"] for s in st)
    for _ in range(3):
        ctx.eval(0.0, tau[100], tau[101], ids, N)
    ms = []
    for _ in range(5):
        ctx.eval(0.0, tau[100], tau[101], ids, N)
        ms.append(ctx.last_device_ms())
    m = float(np.median(ms))
    print("N=%8d orders 0:%d  %.3f ms  %.3e diagram evals/s  alg %.2f TFLOP/s = %.1f%% of measured FP64 peak  (leaves %d, exec FP64 est %.1f%%)"
          % (N, max_order, m, N * tops / (m * 1e-3), flops * N / (m * 1e-3) / 1e12, 100 * flops * N / (m * 1e-3) / 1e12 / peak,
             leaves, 100 * sum(s["n_leaves"] * (4 * (3 * b.order + 1) + 2) * 2 for s, b in zip(st, bold)) * N / (m * 1e-3) / 1e12 / peak))
