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


def bethe_two_orbital(n_tau=128, mu_bethe=0.25, beta=10.0, V=0.5, t_bethe=0.5):
    """test/bethe.jl:61-89: two non-interacting spin orbitals (H = 0), each hybridised with a Bethe bath centred at
    mu_bethe; pairs as the `hybridization=` constructor generates them (src/expansion.jl:239-258)."""
    f = FockSpace([[1], [2]])
    ed = EDCore(f, 0.0 * f.n_op(1))
    grid = ImaginaryTimeGrid(beta, n_tau)
    D = bethe_dos_gf(grid, t=t_bethe, eps=mu_bethe) * V ** 2
    pairs = []
    for o in (1, 2):
        pairs += [InteractionPair(f.c_dag(o), f.c(o), D), InteractionPair(f.c(o), f.c_dag(o), reverse_gf(D))]
    return Expansion(ed, grid, pairs), grid, f


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


def two_site_dimer_exact_rho(beta=1.0, e1=0.5, e2=2.0, V=0.5):
    """Exact occupation probabilities (empty, occupied) of level 1 of the two-site dimer of test/dimers.jl:38-59."""
    w, v = np.linalg.eigh(np.array([[e1, V], [V, e2]]))
    Z = 1 + np.exp(-beta * w).sum() + np.exp(-beta * (e1 + e2))
    occ = ((v[0] ** 2 * np.exp(-beta * w)).sum() + np.exp(-beta * (e1 + e2))) / Z
    return np.array([1 - occ, occ])


def bold_entries(make, orders, n_pts_after_max=None):
    out = []
    for o in orders:
        for k in ([0] if o == 0 else range(1, min(2 * o - 1, n_pts_after_max or 10 ** 9) + 1)):
            out.append((o, k))
    return out


def two_level_mixed(n_tau=12, beta=2.0, theta=0.0):
    """Two degenerate levels with inter-level hybridisation: the one-particle sector is a 2x2 block
    (rotated by `theta` inside the degenerate eigenspace).  Exercises d_s > 1."""
    f = FockSpace([["a"], ["b"]])
    H = 0.3 * (f.n_op("a") + f.n_op("b")) + 0.7 * f.n_op("a") @ f.n_op("b")
    mix = f.c_dag("a") @ f.c("b") + f.c_dag("b") @ f.c("a")
    ed = EDCore(f, H, symmetry_breakers=[mix])
    for s, d in enumerate(ed.dims):
        if d == 2:
            c, sn = np.cos(theta), np.sin(theta)
            ed.unitaries[s] = ed.unitaries[s] @ np.array([[c, -sn], [sn, c]])
    grid = ImaginaryTimeGrid(beta, n_tau)
    D = delta_dos_gf(grid, 0.4) * 0.3
    pairs = []
    for x, y in (("a", "a"), ("b", "b"), ("a", "b"), ("b", "a")):
        pairs += [InteractionPair(f.c_dag(x), f.c(y), D), InteractionPair(f.c(y), f.c_dag(x), ph_conj(D))]
    ex = Expansion(ed, grid, pairs)
    add_corr_operators(ex, (f.c("a"), f.c_dag("a")))
    return ex, grid, f


def three_orbital(n_tau=12, beta=2.0):
    """Three spinless orbitals coupled by hopping: sectors by total particle number with dimensions
    {1, 3, 3, 1}.  Exercises 3x3 blocks (odd column group of the block walker, 3-dimensional edge shapes)."""
    lab = ["a", "b", "c"]
    f = FockSpace([[x] for x in lab])
    n, c, cd = f.n_op, f.c, f.c_dag
    H = 0.2 * n("a") + 0.35 * n("b") - 0.1 * n("c") + 0.5 * (n("a") @ n("b") + n("b") @ n("c")) + 0.8 * n("a") @ n("c")
    H = H + 0.3 * (cd("a") @ c("b") + cd("b") @ c("a")) + 0.45 * (cd("b") @ c("c") + cd("c") @ c("b"))
    ed = EDCore(f, H)
    grid = ImaginaryTimeGrid(beta, n_tau)
    D = delta_dos_gf(grid, [0.4, -0.6], [0.3, 0.2])
    pairs = []
    for x, y in (("a", "a"), ("b", "b"), ("c", "c"), ("a", "c"), ("c", "a")):
        pairs += [InteractionPair(cd(x), c(y), D), InteractionPair(c(y), cd(x), ph_conj(D))]
    ex = Expansion(ed, grid, pairs)
    add_corr_operators(ex, (c("b"), cd("b")))
    return ex, grid, f


def two_band(n_tau=16, beta=8.0, U=2.0, J=0.2, e_k=2.3, big_blocks=False):
    """bench/two_band_eg_model_discrete_bath: two-band e_g model, 16 Fock states, 9 sectors with
    dimensions {1,1,1,1,2,2,2,2,4}, 16 interaction pairs, discrete bath."""
    labels = [[s, o] for s in ("up", "dn") for o in (1, 2)]
    f = FockSpace(labels)
    mu = (3 * U - 5 * J) / 2 - 1.5
    n, c, cd = f.n_op, f.c, f.c_dag
    H = -mu * sum(n("up", o) + n("dn", o) for o in (1, 2)) + U * sum(n("up", o) @ n("dn", o) for o in (1, 2))
    H = H + (U - 2 * J) * sum(n("up", o1) @ n("dn", o2) for o1 in (1, 2) for o2 in (1, 2) if o1 != o2)
    H = H + (U - 3 * J) * sum(n(s, o1) @ n(s, o2) for s in ("up", "dn") for o1 in (1, 2) for o2 in (1, 2) if o2 < o1)
    H = H - J * sum(cd("up", o1) @ cd("dn", o1) @ c("up", o2) @ c("dn", o2) for o1 in (1, 2) for o2 in (1, 2) if o1 != o2)
    H = H - J * sum(cd("up", o1) @ cd("dn", o2) @ c("up", o2) @ c("dn", o1) for o1 in (1, 2) for o2 in (1, 2) if o1 != o2)
    sb = [cd("up", 1) @ c("up", 2) + cd("up", 2) @ c("up", 1), cd("dn", 1) @ c("dn", 2) + cd("dn", 2) @ c("dn", 1)]
    if big_blocks:
        # the same model resolved by particle number only (a spin-flip breaker merges the S_z sectors): sector
        # dimensions {1,4,6,4,1}, i.e. blocks larger than 4x4 (north_star's "sector blocks large enough to be a dense
        # contraction"); physics unchanged, only the bookkeeping of the blocks differs
        sb = sb + [sum(cd("up", o) @ c("dn", o) + cd("dn", o) @ c("up", o) for o in (1, 2))]
    ed = EDCore(f, H, sb)
    grid = ImaginaryTimeGrid(beta, n_tau)
    Delta = delta_dos_gf(grid, [e_k, -e_k], [1.0, 1.0])
    ips = []
    for s in ("up", "dn"):
        for o in (1, 2):
            ips += [InteractionPair(cd(s, o), c(s, o), Delta), InteractionPair(c(s, o), cd(s, o), ph_conj(Delta))]
    for s in ("up", "dn"):
        ips += [InteractionPair(cd(s, 1), c(s, 2), Delta), InteractionPair(c(s, 2), cd(s, 1), ph_conj(Delta)),
                InteractionPair(cd(s, 2), c(s, 1), Delta), InteractionPair(c(s, 1), cd(s, 2), ph_conj(Delta))]
    ex = Expansion(ed, grid, ips)
    add_corr_operators(ex, (c("up", 1), cd("up", 1)))
    return ex, grid, f
