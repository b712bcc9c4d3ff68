"""The reference's bench/bethe_gf_convergence/bethe_gf_convergence.jl on the GPU library: spinless level on a
Bethe bath (beta = 10, t = 2, V = t/2), inchworm! followed by correlator_2p, results stored in the same HDF5
layout (group `data`: attributes beta, ntau, n_pts_after_max, N_samples; datasets orders, orders_bare,
orders_gf, tau, gf, gf_ref), so the plotting script next to the reference's driver reads the file unchanged.

usage: bethe_gf_convergence.py order ntau N_samples [--n_pts_after_max K] [--out FILE]"""
import argparse, hashlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from qinchworm_b200 import h5out, ppgf
from qinchworm_b200.ed import EDCore, FockSpace
from qinchworm_b200.expansion import Expansion, InteractionPair, add_corr_operators
from qinchworm_b200.gf import ImaginaryTimeGrid, bethe_dos_gf, ph_conj
from qinchworm_b200.inchworm import Solver, correlator_2p, inchworm

ap = argparse.ArgumentParser()
ap.add_argument("order", type=int); ap.add_argument("ntau", type=int); ap.add_argument("N_samples", type=int)
ap.add_argument("--n_pts_after_max", type=int, default=None); ap.add_argument("--out", default=None)
a = ap.parse_args()
beta, mu, t_bethe = 10.0, 0.0, 2.0
V = 0.5 * t_bethe
orders = orders_bare = range(0, a.order + 1)
orders_gf = range(0, a.order)

f = FockSpace([["1"]])
ed = EDCore(f, -mu * f.n_op("1"))
grid = ImaginaryTimeGrid(beta, a.ntau)
Delta = bethe_dos_gf(grid, t=t_bethe / 2) * V ** 2
ex = Expansion(ed, grid, [InteractionPair(f.c_dag("1"), f.c("1"), Delta), InteractionPair(f.c("1"), f.c_dag("1"), ph_conj(Delta))])
solver = Solver(ex)
inchworm(ex, grid, orders, orders_bare, a.N_samples, n_pts_after_max=a.n_pts_after_max, solver=solver)
ppgf.normalize(ex)
add_corr_operators(ex, (f.c("1"), f.c_dag("1")))
g = correlator_2p(ex, grid, orders_gf, a.N_samples, solver=solver)[0]
md5 = hashlib.md5(np.asarray(g, dtype="<c16").tobytes()).hexdigest()
out = a.out or "data_order_%s_ntau_%d_N_samples_%d_md5_%s.h5" % ("%d:%d" % (0, a.order), a.ntau, a.N_samples, md5)
h5out.bethe_gf_results(out, beta, a.ntau, a.N_samples, orders, orders_bare, orders_gf, grid.tau, g,
                       -np.asarray(Delta.data).ravel() / V ** 2, n_pts_after_max=a.n_pts_after_max)
print("filename =", out)
print("max |gf - gf_ref| = %.3e" % np.abs(g + np.asarray(Delta.data).ravel() / V ** 2).max())
