"""Small run of every kernel for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool racecheck python profiles/sanitize.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import models
from qinchworm_b200 import lib
from qinchworm_b200.inchworm import Solver, correlator_2p, inchworm
what = sys.argv[1:] or ["scalar", "block", "mma"]
if "scalar" in what:
    ex, grid, f = models.anderson(n_tau=10, corr=True)
    ctx = lib.Context(device=0)
    solver = Solver(ex, ctx=ctx)
    inchworm(ex, grid, range(0, 4), range(0, 4), 2 ** 7, solver=solver)                       # step kernel (bare) + run kernel
    inchworm(ex, grid, range(0, 3), range(0, 3), 2 ** 7, solver=solver, device_resident=False)  # step kernel per step + scale_P
    g = correlator_2p(ex, grid, range(0, 3), 2 ** 6, solver=solver)[0]                        # batched (gridDim.z)
    print("scalar ok", float(np.abs(ex.P).sum()), float(np.abs(g).sum()))
    ctx.close()
if "block" in what:
    ex, grid, f = models.two_band(n_tau=6)
    ctx = lib.Context(device=0)
    inchworm(ex, grid, range(0, 3), range(0, 3), 2 ** 6, solver=Solver(ex, ctx=ctx))          # block walker, fused tail
    print("block ok", float(np.abs(ex.P).sum()))
    ctx.close()
if "mma" in what:
    ex, grid, f = models.two_band(n_tau=6, big_blocks=True)
    ctx = lib.Context(device=0)
    inchworm(ex, grid, range(0, 3), range(0, 3), 2 ** 5, solver=Solver(ex, ctx=ctx))          # DMMA kernel
    print("mma ok", float(np.abs(ex.P).sum()))
    ctx.close()
