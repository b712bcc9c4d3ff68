"""Profiling driver: README Anderson configuration (C1), `runs` device-resident inchworm! runs.
Used under ncu (profiles/run_profiles.sh); numbers printed by a run under ncu are not bench values."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import models  # noqa: E402
from qinchworm_b200 import lib  # noqa: E402
from qinchworm_b200.inchworm import MODE_BARE, Solver, _bold_entries  # noqa: E402

n_tau = int(sys.argv[1]) if len(sys.argv) > 1 else 200
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
runs = int(sys.argv[3]) if len(sys.argv) > 3 else 1
max_order = int(sys.argv[4]) if len(sys.argv) > 4 else 4
ex, grid, f = models.anderson(n_tau=n_tau)
ctx = lib.Context(device=0)
solver = Solver(ex, ctx=ctx)
orders = range(0, max_order + 1)
bare = [solver.make_entry(MODE_BARE, o, 2 * o, N) for o in orders]
bold = _bold_entries(solver, orders, N, None, None)
P0 = ex.P.copy()
for _ in range(runs):
    ctx.set_P(0, P0)
    ctx.inchworm_run([t.entry_id for t in bare], [t.entry_id for t in bold], N, want_contribs=False)
    print("device ms", ctx.last_device_ms(), "launches", ctx.launch_count())
