"""Per-step state update of the bold propagators — host mirror of src/ppgf.jl.

The P table is `expansion.P[k, packed blocks]` (grid point k, sector blocks concatenated,
column-major inside a block).  These O(n_tau) updates stay on the host in the step-level API
(SURVEY §8b); the run-level API of the library performs the same update on the device.
"""
from __future__ import annotations

import math

import numpy as np

__all__ = ["set_ppgf", "normalize_at", "normalize", "partition_function", "density_matrix"]


def set_ppgf(expansion, k_f: int, value):
    """set_ppgf!(P, tau_i, tau_f, val): P(tau_f) <- val (src/ppgf.jl:474-504)."""
    expansion.P[k_f, :] = np.asarray(value, dtype=complex)


def _diag_indices(expansion):
    key = (tuple(expansion.dims), tuple(expansion.boff))
    cached = getattr(expansion, "_diag_cache", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    idx = []
    for s, d in enumerate(expansion.dims):
        idx += [expansion.boff[s] + i + d * i for i in range(d)]
    idx = np.asarray(idx, dtype=int)
    try:
        expansion._diag_cache = (key, idx)
    except AttributeError:
        pass
    return idx


def normalize_at(expansion, k_f: int):
    """normalize!(P, tau): lambda = log(max_s max diag(-Im P_s(tau))) / tau, then every stored grid
    value is multiplied by exp(-lambda tau_k) (src/ppgf.jl:646-668).  Returns lambda."""
    tau = expansion.grid.tau
    p_max = -float(expansion.P[k_f].imag[_diag_indices(expansion)].min())
    lam = (math.log(p_max) if p_max > 0.0 else float("nan")) / float(tau[k_f])
    expansion.P *= np.exp(tau * -lam)[:, None]
    return lam


def partition_function(expansion) -> complex:
    """Z = sum_s i tr P_s(beta) (src/ppgf.jl:611-617)."""
    return complex(1j * expansion.P[-1, _diag_indices(expansion)].sum())


def normalize(expansion, beta: float | None = None):
    """normalize!(P, beta): lambda = log(Z)/beta, P *= exp(-lambda tau) (src/ppgf.jl:628-635)."""
    beta = expansion.grid.beta if beta is None else beta
    lam = np.log(partition_function(expansion)) / beta
    expansion.P *= np.exp(-expansion.grid.tau * lam)[:, None]
    return lam


def density_matrix(expansion):
    """rho = i P(beta) as a list of blocks (src/ppgf.jl:727-731)."""
    return [1j * expansion.block(expansion.P[-1], s) for s in range(expansion.S)]
