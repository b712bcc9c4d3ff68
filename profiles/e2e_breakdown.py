import os, sys, time
ROOT='/root/repo'
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import models
from qinchworm_b200 import lib
from qinchworm_b200.inchworm import MODE_BARE, Solver, _bold_entries, inchworm
ex, grid, f = models.anderson(n_tau=200)
ctx = lib.Context(device=0)
solver = Solver(ex, ctx=ctx)
P0 = ex.P.copy()
N = 1024
for rep in range(4):
    ex.P[:] = P0
    t0 = time.perf_counter()
    bare = [solver.make_entry(MODE_BARE, o, 2 * o, N) for o in range(5)]
    bold = _bold_entries(solver, range(5), N, None, None)
    t1 = time.perf_counter()
    solver.upload_P()
    t2 = time.perf_counter()
    hist = ctx.inchworm_run([td.entry_id for td in bare], [td.entry_id for td in bold], N)
    t3 = time.perf_counter()
    P = ctx.get_P()
    t4 = time.perf_counter()
    ord_of = np.array([td.order for td in bare + bold])
    Po = {o: hist[:, ord_of == o, :].sum(axis=1) for o in range(5)}
    t5 = time.perf_counter()
print("entries %.0f us, upload_P %.0f us, inchworm_run %.0f us (device %.0f us), get_P %.0f us, order sums %.0f us"
      % ((t1-t0)*1e6, (t2-t1)*1e6, (t3-t2)*1e6, ctx.last_device_ms()*1e3, (t4-t3)*1e6, (t5-t4)*1e6))
t = time.perf_counter(); h2 = ctx.inchworm_run([td.entry_id for td in bare], [td.entry_id for td in bold], N, want_contribs=False); print("run without contribs %.0f us" % ((time.perf_counter()-t)*1e6))
ts = []
for rep in range(20):
    ex.P[:] = P0
    t = time.perf_counter()
    inchworm(ex, grid, range(5), range(5), N, solver=solver)
    ts.append((time.perf_counter() - t) * 1e6)
print("inchworm(...) public API: min %.0f us, median %.0f us, max %.0f us (device %.0f us)" % (min(ts), float(np.median(ts)), max(ts), ctx.last_device_ms() * 1e3))
