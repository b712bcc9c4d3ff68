"""Synthetic model builders shared by the tests and bench.py (inputs only, no hot-path code)."""
import numpy as np

from qinchworm_b200.ed import EDCore, FockSpace
from qinchworm_b200.expansion import Expansion, InteractionPair, add_corr_operators
from qinchworm_b200.gf import (ImaginaryTimeGF, ImaginaryTimeGrid, SplineInterpolatedGF, bethe_dos_gf,
                                delta_dos_gf, ph_conj, reverse_gf)


def single_level(n_tau=20, beta=10.0, mu=0.1, eps=0.1, V=-0.1, spline=True, rev="reverse"):
    """The model of test/inchworm.jl:55-66 and test/topology_eval.jl:40-66: H = -mu n."""
    f = FockSpace([["0"]])
    ed = EDCore(f, -mu * f.n_op("0"))
    grid = ImaginaryTimeGrid(beta, n_tau)
    D = delta_dos_gf(grid, eps) * V ** 2
    if rev == "reverse":        # (t1, t2) -> -Delta[t2, t1, false]   test/inchworm.jl:187
        Drev = reverse_gf(D)
    else:                       # (t1, t2) -> Delta[t2, t1]           test/topology_eval.jl:61
        Drev = ImaginaryTimeGF(grid, np.concatenate([[D.data[0]], -D.data[::-1][1:]]))
    wrap = SplineInterpolatedGF if spline else (lambda g: g)
    ex = Expansion(ed, grid, [InteractionPair(f.c_dag("0"), f.c("0"), wrap(D)),
                              InteractionPair(f.c("0"), f.c_dag("0"), wrap(Drev))],
                   interpolate_ppgf=spline)
    return ex, grid, f


def anderson(n_tau=200, beta=10.0, eps=0.1, U=1.0, D=2.0, V=0.5, corr=False):
    """README.md:34-158: single-orbital Anderson model, semi-elliptic (Bethe) bath."""
    f = FockSpace([["up"], ["dn"]])
    H = eps * (f.n_op("up") + f.n_op("dn")) + U * f.n_op("up") @ f.n_op("dn")
    ed = EDCore(f, H)
    grid = ImaginaryTimeGrid(beta, n_tau)
    Delta = bethe_dos_gf(grid, t=D / 2) * V ** 2
    pairs = [InteractionPair(f.c_dag("up"), f.c("up"), Delta),
             InteractionPair(f.c("up"), f.c_dag("up"), ph_conj(Delta)),
             InteractionPair(f.c_dag("dn"), f.c("dn"), Delta),
             InteractionPair(f.c("dn"), f.c_dag("dn"), ph_conj(Delta))]
    ex = Expansion(ed, grid, pairs)
    if corr:
        add_corr_operators(ex, (-f.c("up"), f.c_dag("up")))
        add_corr_operators(ex, (f.n_op("up"), f.n_op("dn")))
        add_corr_operators(ex, (f.c_dag("up") @ f.c("dn"), f.c_dag("dn") @ f.c("up")))
    return ex, grid, f


def bethe_two_state(n_tau=64, beta=8.0, V=1.0, t_bethe=1.0, mu=0.0):
    """bench/bethe_gf_convergence: spinless level on a Bethe bath (2 sectors), G = <c(tau) c^dag(0)>."""
    f = FockSpace([["0"]])
    ed = EDCore(f, -mu * f.n_op("0"))
    grid = ImaginaryTimeGrid(beta, n_tau)
    Delta = bethe_dos_gf(grid, t=t_bethe) * V ** 2
    ex = Expansion(ed, grid, [InteractionPair(f.c_dag("0"), f.c("0"), Delta),
                              InteractionPair(f.c("0"), f.c_dag("0"), ph_conj(Delta))])
    add_corr_operators(ex, (f.c("0"), f.c_dag("0")))
    return ex, grid, f


def hubbard_dimer_impurity(n_tau=32, beta=1.0, U=4.0, eps2=0.0, V=0.5):
    """test/dimers.jl:124-178: two-orbital impurity (modes 1, 2), each hybridised with one bath level."""
    f = FockSpace([[1], [2]])
    eps1 = -0.5 * U
    H = U * f.n_op(1) @ f.n_op(2) + eps1 * (f.n_op(1) + f.n_op(2))
    ed = EDCore(f, H)
    grid = ImaginaryTimeGrid(beta, n_tau)
    D1 = delta_dos_gf(grid, eps2) * V ** 2
    pairs = []
    for o in (1, 2):
        pairs += [InteractionPair(f.c_dag(o), f.c(o), D1), InteractionPair(f.c(o), f.c_dag(o), reverse_gf(D1))]
    return Expansion(ed, grid, pairs), grid, f


def hubbard_dimer_exact_rho(beta=1.0, U=4.0, eps2=0.0, V=0.5):
    """Exact reduced density matrix (diagonal in the impurity Fock basis) of the 4-mode dimer."""
    f = FockSpace([[1], [2], [3], [4]])
    eps1 = -0.5 * U
    H = (U * f.n_op(1) @ f.n_op(2) + eps1 * (f.n_op(1) + f.n_op(2)) + eps2 * (f.n_op(3) + f.n_op(4))
         + V * (f.c_dag(1) @ f.c(3) + f.c_dag(3) @ f.c(1)) + V * (f.c_dag(2) @ f.c(4) + f.c_dag(4) @ f.c(2)))
    w, v = np.linalg.eigh(H)
    rho = (v * np.exp(-beta * (w - w.min()))) @ v.T
    rho /= np.trace(rho)
    red = np.zeros((4, 4))
    for a in range(16):
        for b in range(16):
            if (a >> 2) == (b >> 2):
                red[a & 3, b & 3] += rho[a, b]
    return red


def bold_entries(make, orders, n_pts_after_max=None):
    out = []
    for o in orders:
        for k in ([0] if o == 0 else range(1, min(2 * o - 1, n_pts_after_max or 10 ** 9) + 1)):
            out.append((o, k))
    return out
