"""Diagnostics: per-CTA timeline of one bold step kernel (QIW_TRACE) on the C1 configuration."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import models
from qinchworm_b200 import lib
from qinchworm_b200.inchworm import MODE_BARE, Solver, _bold_entries
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ex, grid, f = models.anderson(n_tau=200)
ctx = lib.Context(device=0)
solver = Solver(ex, ctx=ctx)
max_order = int(sys.argv[2]) if len(sys.argv) > 2 else 4
bold = _bold_entries(solver, range(0, max_order + 1), N, None, None)
ids = [t.entry_id for t in bold]
tau = grid.tau
for _ in range(5):
    ctx.eval(0.0, tau[100], tau[101], ids, N)
os.environ["QIW_TRACE"] = os.path.join(ROOT, "gpurun_out", "trace.csv")
ctx.eval(0.0, tau[100], tau[101], ids, N)
print("device ms", ctx.last_device_ms())
del os.environ["QIW_TRACE"]
t = np.loadtxt(os.path.join(ROOT, "gpurun_out", "trace.csv"), delimiter=",", skiprows=1)
clk = 1.965e3  # cycles per us at max clock
t = np.atleast_2d(t)
t0 = t[:, 8].min()
print("CTAs", len(t), "start spread us", (t[:, 8].max() - t0) / 1e3)
m = t[:, 2] > 0
print("setup us: mean %.2f max %.2f" % (((t[m, 2] - t[m, 1]) / clk).mean(), ((t[m, 2] - t[m, 1]) / clk).max()))
print("  of which roots %.2f times %.2f fill %.2f segments %.2f (means)" % (((t[m, 9] - t[m, 1]) / clk).mean(), ((t[m, 10] - t[m, 9]) / clk).mean(),
      ((t[m, 11] - t[m, 10]) / clk).mean(), ((t[m, 2] - t[m, 11]) / clk).mean()))
print("kernel span us (first CTA start .. last CTA end, globaltimer + cycles): %.2f" % ((t[m, 8] - t0) / 1e3 + (t[m, 4] - t[m, 1]) / clk).max())
print("walk  us: mean %.2f max %.2f" % (((t[m, 3] - t[m, 2]) / clk).mean(), ((t[m, 3] - t[m, 2]) / clk).max()))
print("tail  us: mean %.2f max %.2f" % (((t[m, 4] - t[m, 3]) / clk).mean(), ((t[m, 4] - t[m, 3]) / clk).max()))
lt = t[:, 12] > 0
if lt.any():
    print("last CTA: walk end -> tail end %.2f us; its start offset %.2f us; kernel span incl. tail %.2f us" % (
        ((t[lt, 12] - t[lt, 4]) / clk).max(), ((t[lt, 8] - t0) / 1e3).max(), ((t[lt, 8] - t0) / 1e3 + (t[lt, 12] - t[lt, 1]) / clk).max()))
print("start offsets us: pctl", np.percentile((t[:, 8] - t0) / 1e3, [0, 25, 50, 75, 90, 100]))
for e in sorted(set(t[:, 6])):
    k = (t[:, 6] == e) & m
    if k.sum() == 0:
        continue
    print("entry %d: n %d groups/warp0 %d roots %.2f times %.2f fill %.2f seg %.2f | setup %.2f walk %.2f tail %.2f total %.2f start %.1f" % (
        e, k.sum(), t[k, 7].mean(), ((t[k, 9] - t[k, 1]) / clk).mean(), ((t[k, 10] - t[k, 9]) / clk).mean(),
        ((t[k, 11] - t[k, 10]) / clk).mean(), ((t[k, 2] - t[k, 11]) / clk).mean(),
        ((t[k, 2] - t[k, 1]) / clk).mean(), ((t[k, 3] - t[k, 2]) / clk).mean(),
        ((t[k, 4] - t[k, 3]) / clk).mean(), ((t[k, 4] - t[k, 1]) / clk).mean(), ((t[k, 8] - t0) / 1e3).mean()))
# per-SM busy time
import collections
sm = collections.defaultdict(float)
for row in t[m]:
    sm[int(row[5])] += (row[4] - row[1]) / clk
v = np.array(list(sm.values()))
print("per-SM summed CTA time us: min %.1f mean %.1f max %.1f (n SMs %d)" % (v.min(), v.mean(), v.max(), len(v)))
