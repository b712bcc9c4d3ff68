"""Sector blocks of 5 to 8 rows: the FP64 tensor-core kernel (block_mma_kernel, mma.sync.m8n8k4.f64, warp = sample)
against the general FMA kernel (block_step_kernel<8>, thread = sample, complex arithmetic; QIW_FORCE_COMPLEX=1) on the
two-band model resolved by particle number only (blocks {1,4,6,4,1}), one bold step, orders 0:max_order."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import models
from qinchworm_b200 import lib
from qinchworm_b200.inchworm import Solver, _bold_entries
max_order = int(sys.argv[1]) if len(sys.argv) > 1 else 2
Ns = [int(x) for x in sys.argv[2:]] or [2 ** 8, 2 ** 10, 2 ** 12]
ex, grid, f = models.two_band(n_tau=32, big_blocks=True)
for mode in ("mma", "fma"):
    os.environ["QIW_FORCE_COMPLEX"] = "0" if mode == "mma" else "1"
    ctx = lib.Context(device=0)
    peak = ctx.measure_fp64_peak()
    solver = Solver(ex, ctx=ctx)
    for N in Ns:
        bold = _bold_entries(solver, range(0, max_order + 1), N, None, None)
        ids = [t.entry_id for t in bold]
        st = [ctx.entry_stats(i) for i in ids]
        flops = sum(s["flops_per_sample"] for s in st)
        tops = sum(s["n_top"] for s in st)
        ctx.eval(0.0, grid.tau[15], grid.tau[16], ids, N)
        ms = []
        for _ in range(3):
            r = ctx.eval(0.0, grid.tau[15], grid.tau[16], ids, N)
            ms.append(ctx.last_device_ms())
        m_ = float(np.median(ms))
        print("%s N=%6d orders 0:%d  %.3f ms  %.3e diagram evals/s  alg %.2f TFLOP/s = %.1f%% of measured FP64 peak  checksum %.12e"
              % (mode, N, max_order, m_, N * tops / (m_ * 1e-3), flops * N / (m_ * 1e-3) / 1e12, 100 * flops * N / (m_ * 1e-3) / 1e12 / peak,
                 float(np.abs(r).sum())))
    ctx.close()
