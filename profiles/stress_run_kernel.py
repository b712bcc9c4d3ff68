"""Robustness of the persistent run kernel (cooperative launch, spin grid barrier): many back-to-back runs at several
sample counts (one job per CTA / several jobs per CTA), every run bit-identical to the first."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import models
from qinchworm_b200 import lib
from qinchworm_b200.inchworm import MODE_BARE, Solver, _bold_entries
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ex, grid, f = models.anderson(n_tau=200)
ctx = lib.Context(device=0)
solver = Solver(ex, ctx=ctx)
P0 = ex.P.copy()
for N in (2 ** 10, 2 ** 11, 2 ** 12, 2 ** 9, 2 ** 6):
    bare = [solver.make_entry(MODE_BARE, o, 2 * o, N) for o in range(5)]
    bold = _bold_entries(solver, range(5), N, None, None)
    first, ms = None, []
    t = time.perf_counter()
    for r in range(reps if N == 2 ** 10 else max(reps // 10, 5)):
        ctx.set_P(0, P0)
        l0 = ctx.launch_count()
        ctx.inchworm_run([t_.entry_id for t_ in bare], [t_.entry_id for t_ in bold], N, want_contribs=False)
        ms.append(ctx.last_device_ms())
        P = ctx.get_P()
        if first is None:
            first, nl = P, ctx.launch_count() - l0
        assert np.array_equal(P, first), ("run %d differs" % r)
    print("N=%5d: %d runs, %d launches per run, device ms min %.3f median %.3f max %.3f, all bit-identical (%.1f s)"
          % (N, len(ms), nl, min(ms), float(np.median(ms)), max(ms), time.perf_counter() - t))
