"""Per-SM view of a run-kernel trace (gpurun_out/trace_run.csv, written by profiles/trace_run.py with the `make trace`
build): when each SM's CTAs have delivered their partial rows in the middle step, grouped by the pair of entries the
SM hosts.  The step's grid barrier opens when the slowest SM is done.
usage: python profiles/trace_by_sm.py [trace_run.csv]"""
import collections, csv, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "trace_run.csv")
rows = list(csv.DictReader(open(path)))
khz = 1.965e3          # cycles per microsecond
sm = collections.defaultdict(list)
for r in rows:
    sm[int(r["smid"])].append(r)
res = []
for s, g in sm.items():
    done = max(float(x["jobs_done"]) for x in g) / khz
    pair = "+".join(sorted("e%s%s" % (x["entry"], "" if x["n_jobs"] == "1" else "x" + x["n_jobs"]) for x in g))
    res.append((done, pair))
res.sort()
d = np.array([r[0] for r in res])
print("%d SMs; partial rows delivered after (us since the step's start): min %.2f  median %.2f  p90 %.2f  max %.2f"
      % (len(res), d.min(), np.median(d), np.percentile(d, 90), d.max()))
by = collections.defaultdict(list)
for t, p in res:
    by[p].append(t)
print("%-16s %5s %8s %8s" % ("entries on the SM", "SMs", "median", "max"))
for k, v in sorted(by.items(), key=lambda kv: -np.median(kv[1])):
    print("%-16s %5d %8.2f %8.2f" % (k, len(v), np.median(v), max(v)))
