"""C1 device-resident run: persistent run kernel vs one step kernel per step (QIW_NO_RUN_KERNEL=1), same context —
device ms per run and the largest relative difference of the final P tables."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import models  # noqa: E402
from qinchworm_b200 import lib  # noqa: E402
from qinchworm_b200.inchworm import MODE_BARE, Solver, _bold_entries  # noqa: E402

n_tau = int(sys.argv[1]) if len(sys.argv) > 1 else 200
Ns = [int(x) for x in sys.argv[2:]] or [1024]
ex, grid, f = models.anderson(n_tau=n_tau)
ctx = lib.Context(device=0)
solver = Solver(ex, ctx=ctx)
P0 = ex.P.copy()
for N in Ns:
    bare = [solver.make_entry(MODE_BARE, o, 2 * o, N) for o in range(5)]
    bold = _bold_entries(solver, range(5), N, None, None)
    res = {}
    for mode in ("run", "step"):
        os.environ["QIW_NO_RUN_KERNEL"] = "0" if mode == "run" else "1"
        ms = []
        for _ in range(6):
            ctx.set_P(0, P0)
            l0 = ctx.launch_count()
            hist = ctx.inchworm_run([t.entry_id for t in bare], [t.entry_id for t in bold], N, want_contribs=True)
            ms.append(ctx.last_device_ms())
            nl = ctx.launch_count() - l0
        res[mode] = (ctx.get_P(), hist, float(np.median(ms[1:])), nl)
    d = np.abs(res["run"][0] - res["step"][0]).max() / np.abs(res["step"][0]).max()
    dh = np.abs(res["run"][1] - res["step"][1]).max() / np.abs(res["step"][1]).max()
    print("N=%d: run kernel %.3f ms (%d launches), per-step launches %.3f ms (%d launches); max rel diff P %.2e, contributions %.2e"
          % (N, res["run"][2], res["run"][3], res["step"][2], res["step"][3], d, dh))

# the step seam by itself: 198 x (qiw_eval + qiw_scale_P) with no host bookkeeping — the floor of the host-stepped loop
import time
N = Ns[0]
bold = _bold_entries(solver, range(5), N, None, None)
ids = [t.entry_id for t in bold]
tau = grid.tau
ctx.set_P(0, P0)
row = P0[5].copy()
for rep in range(3):
    t = time.perf_counter()
    gpu = 0.0
    for n in range(1, n_tau - 1):
        r = ctx.eval(0.0, tau[n], tau[n + 1], ids, N)
        gpu += ctx.last_device_ms()
        ctx.scale_P(n + 1, row, 0.0)
    dt = (time.perf_counter() - t) * 1e3
print("step seam alone: %d x (qiw_eval + qiw_scale_P) %.2f ms wall (%.1f us per step), of which kernel time %.2f ms" % (n_tau - 2, dt, dt * 1e3 / (n_tau - 2), gpu))
from qinchworm_b200.inchworm import inchworm
ex.P[:] = P0
inchworm(ex, grid, range(5), range(5), N, solver=solver, device_resident=False)
ex.P[:] = P0
t = time.perf_counter()
inchworm(ex, grid, range(5), range(5), N, solver=solver, device_resident=False)
print("inchworm(..., device_resident=False): %.2f ms wall" % ((time.perf_counter() - t) * 1e3))
