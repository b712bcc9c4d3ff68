"""Throughput of one bold step of the two-band e_g model (C4: S=9 sectors, blocks up to 4x4, 16 pairs)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import models
from qinchworm_b200 import lib
from qinchworm_b200.inchworm import Solver, _bold_entries
max_order = int(sys.argv[1]) if len(sys.argv) > 1 else 3
Ns = [int(x) for x in sys.argv[2:]] or [2 ** 10]
ex, grid, f = models.two_band(n_tau=64)
ctx = lib.Context(device=0)
peak = ctx.measure_fp64_peak()
solver = Solver(ex, ctx=ctx)
tau = grid.tau
import time
t0 = time.perf_counter()
bold = _bold_entries(solver, range(0, max_order + 1), 1024, None, None)
print("compile %.2f s" % (time.perf_counter() - t0))
ids = [t.entry_id for t in bold]
st = [ctx.entry_stats(i) for i in ids]
flops = sum(s["flops_per_sample"] for s in st)
tops = sum(s["n_top"] for s in st)
leaves = sum(s["n_leaves"] for s in st)
print("topologies %d leaves %d edges %d chain flops/sample %.3e" % (tops, leaves, sum(s["n_edges"] for s in st), flops))
for N in Ns:
    for _ in range(2):
        ctx.eval(0.0, tau[30], tau[31], ids, N)
    ms = []
    for _ in range(3):
        ctx.eval(0.0, tau[30], tau[31], ids, N)
        ms.append(ctx.last_device_ms())
    m = float(np.median(ms))
    print("N=%7d orders 0:%d  %.3f ms  %.3e diagram evals/s  alg %.3f TFLOP/s = %.2f%% of measured FP64 peak %.1f"
          % (N, max_order, m, N * tops / (m * 1e-3), flops * N / (m * 1e-3) / 1e12, 100 * flops * N / (m * 1e-3) / 1e12 / peak, peak))
