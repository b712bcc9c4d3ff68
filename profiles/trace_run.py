"""Per-CTA phase timeline of the middle step of the persistent run kernel (needs `make trace`).
usage: QIW_LIB=qinchworm.jl_b200/libqinchworm_cuda_trace.so python profiles/trace_run.py [n_tau] [N]"""
import csv, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ.setdefault("QIW_LIB", os.path.join(ROOT, "qinchworm.jl_b200", "libqinchworm_cuda_trace.so"))
import numpy as np
import models
from qinchworm_b200 import lib
from qinchworm_b200.inchworm import MODE_BARE, Solver, _bold_entries
n_tau = int(sys.argv[1]) if len(sys.argv) > 1 else 200
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
ex, grid, f = models.anderson(n_tau=n_tau)
ctx = lib.Context(device=0)
solver = Solver(ex, ctx=ctx)
bare = [solver.make_entry(MODE_BARE, o, 2 * o, N) for o in range(5)]
bold = _bold_entries(solver, range(5), N, None, None)
P0 = ex.P.copy()
path = os.path.join(ROOT, "gpurun_out", "trace_run.csv")
for k in range(3):
    if k == 2:
        os.environ["QIW_TRACE"] = path
    ctx.set_P(0, P0)
    ctx.inchworm_run([t.entry_id for t in bare], [t.entry_id for t in bold], N, want_contribs=False)
    print("device ms", ctx.last_device_ms())
rows = list(csv.DictReader(open(path)))
ghz = 1.965
info = {t.entry_id: (t.order, t.n_pts_after) for t in bold}
groups = collections.defaultdict(list)
for r in rows:
    groups[(int(r["entry"]), int(r["n_jobs"]), int(r["n_sub"]))].append(r)
cols = ["times", "fill", "seg", "walk", "jobs_done", "barrier", "reduce", "update"]
print("cycles -> us at %.3f GHz; cumulative since the step's start, median over the group's CTAs (max in brackets for jobs_done)" % ghz)
print("%-22s %4s " % ("entry(order,k) jobs m", "ctas") + " ".join("%9s" % c for c in cols))
for key in sorted(groups, key=lambda k: info.get(k[0], (9, 9))):
    g = groups[key]
    med = [np.median([float(r[c]) for r in g]) / ghz / 1e3 for c in cols]
    mx = max(float(r["jobs_done"]) for r in g) / ghz / 1e3
    print("%-22s %4d " % ("%s x%d m%d" % (info.get(key[0]), key[1], key[2]), len(g)) + " ".join("%9.2f" % v for v in med) + "  [%.2f]" % mx)
ends = np.array([float(r["end_ns"]) for r in rows])
print("spread of the step's end over CTAs (globaltimer): %.2f us" % ((ends.max() - ends.min()) / 1e3))
