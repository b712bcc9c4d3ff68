"""Imaginary-time grid and scalar Green's-function containers (host side).

Stand-ins for the pieces of Keldysh.jl the hot path touches: `ImaginaryTimeGrid`,
scalar `ImaginaryTimeGF` built from a density of states, `ph_conj` (src/utility.jl:210-217) and the
`SplineInterpolatedGF` wrapper (src/spline_gf.jl:65-88).  Values follow Keldysh.jl's convention
for a fermionic GF on the Matsubara branch: G(tau) = -i * int dw rho(w) exp(-w tau) / (1 + exp(-beta w)).
"""
from __future__ import annotations

import numpy as np

__all__ = ["ImaginaryTimeGrid", "ImaginaryTimeGF", "SplineInterpolatedGF", "delta_dos_gf",
           "bethe_dos_gf", "ph_conj", "reverse_gf"]

GRID_BILINEAR = 0   # Keldysh.jl generic grid interpolation
CUBIC_SPLINE = 1    # natural cubic spline in tau_f - tau_i (src/spline_gf.jl:193-219)


class ImaginaryTimeGrid:
    def __init__(self, beta: float, n_tau: int):
        self.beta = float(beta)
        self.n_tau = int(n_tau)
        self.tau = np.linspace(0.0, self.beta, self.n_tau)

    def __len__(self):
        return self.n_tau

    def __getitem__(self, k):
        return self.tau[k]


class ImaginaryTimeGF:
    """Scalar imaginary-time function stored on the grid: data[k] = G(tau_k, 0)."""
    kind = GRID_BILINEAR

    def __init__(self, grid: ImaginaryTimeGrid, data):
        self.grid = grid
        self.data = np.asarray(data, dtype=complex).copy()
        assert self.data.shape == (grid.n_tau,)

    def __mul__(self, x):
        return type(self)(self.grid, self.data * x)

    __rmul__ = __mul__


class SplineInterpolatedGF(ImaginaryTimeGF):
    """Marks a GF to be evaluated by its natural cubic spline (src/spline_gf.jl:65-88)."""
    kind = CUBIC_SPLINE

    def __init__(self, gf_or_grid, data=None):
        if isinstance(gf_or_grid, ImaginaryTimeGF):
            super().__init__(gf_or_grid.grid, gf_or_grid.data)
        else:
            super().__init__(gf_or_grid, data)


def delta_dos_gf(grid, eps, weights=None):
    """kd.ImaginaryTimeGF(kd.DeltaDOS(eps[, weights]), grid)."""
    eps = np.atleast_1d(np.asarray(eps, dtype=float))
    w = np.ones_like(eps) if weights is None else np.atleast_1d(np.asarray(weights, dtype=float))
    tau = grid.tau[:, None]
    beta = grid.beta
    # exp(-e tau) / (1 + exp(-beta e)), written to stay finite for negative e
    e = eps[None, :]
    val = np.where(e >= 0, np.exp(-e * tau) / (1 + np.exp(-beta * e)),
                   np.exp(e * (beta - tau)) / (1 + np.exp(beta * e)))
    return ImaginaryTimeGF(grid, -1j * (val * w[None, :]).sum(axis=1))


def bethe_dos_gf(grid, t=1.0, eps=0.0, n_quad=2000):
    """kd.ImaginaryTimeGF(kd.bethe_dos(t=t, ϵ=eps), grid): semicircle of half-width 2t.

    rho(w) = sqrt(4t^2 - (w-eps)^2) / (2 pi t^2); Gauss-Chebyshev (2nd kind) quadrature, which
    integrates the square-root weight exactly.
    """
    D = 2.0 * t
    k = np.arange(1, n_quad + 1)
    x = np.cos(k * np.pi / (n_quad + 1))
    wq = np.pi / (n_quad + 1) * np.sin(k * np.pi / (n_quad + 1)) ** 2  # int sqrt(1-x^2) f(x) dx
    w = eps + D * x
    beta = grid.beta
    tau = grid.tau[:, None]
    ww = w[None, :]
    f = np.where(ww >= 0, np.exp(-ww * tau) / (1 + np.exp(-beta * ww)),
                 np.exp(ww * (beta - tau)) / (1 + np.exp(beta * ww)))
    # rho(w) dw = (2/pi) sqrt(1-x^2) dx
    val = (2.0 / np.pi) * (f * wq[None, :]).sum(axis=1)
    return ImaginaryTimeGF(grid, -1j * val)


def ph_conj(g):
    """g(tau) -> g(beta - tau)  (src/utility.jl:210-217)."""
    return type(g)(g.grid, g.data[::-1])


def reverse_gf(g):
    """`(t1, t2) -> -g[t2, t1, false]` of the reference's tests (test/inchworm.jl:187): for a
    fermionic Matsubara GF this is the same reversed array as `ph_conj`."""
    return type(g)(g.grid, g.data[::-1])
