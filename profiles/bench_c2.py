"""C2 (bench/bethe_gf_convergence): correlator_2p G(tau) of a spinless level on a Bethe bath, orders_gf 0:3,
N_samples sweep; all grid points in one qiw_eval_batch launch vs one qiw_eval per point."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import models
from qinchworm_b200 import lib
from qinchworm_b200.inchworm import Solver, correlator_2p, inchworm
n_tau = int(sys.argv[1]) if len(sys.argv) > 1 else 128
Ns = [int(x) for x in sys.argv[2:]] or [2 ** 10, 2 ** 13, 2 ** 16]
ex, grid, f = models.bethe_two_state(n_tau=n_tau)
ctx = lib.Context(device=0)
solver = Solver(ex, ctx=ctx)
inchworm(ex, grid, range(0, 4), range(0, 4), 2 ** 10, solver=solver)
tops = sum(len(lib.topologies(o, k, True)[1]) for o in range(0, 4) for k in ([0] if o == 0 else range(1, 2 * o)))
for N in Ns:
    for batch in (True, False):
        correlator_2p(ex, grid, range(0, 4), N, solver=solver, batch=batch)
        t = time.perf_counter()
        g = correlator_2p(ex, grid, range(0, 4), N, solver=solver, batch=batch)[0]
        dt = time.perf_counter() - t
        print("N=%7d n_tau=%d %s: %.2f ms wall, %.3e diagram evals/s" % (N, n_tau, "batched" if batch else "stepped", dt * 1e3,
              N * tops * (n_tau - 1) / dt))
