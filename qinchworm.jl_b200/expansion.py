"""`Expansion` / `InteractionPair` — host mirror of src/expansion.jl:60-312.

An `Expansion` bundles what the hot path reads: sector dimensions, atomic energies (for the bare
propagator P0), the bold propagator table P on the imaginary-time grid, the interaction pairs
(operator_i, operator_f, scalar propagator Delta) and the correlator operator pairs.
`flatten()` produces the plain arrays that cross the C ABI in `qiw_set_model` / `qiw_set_delta`.
"""
from __future__ import annotations

import numpy as np

from .ed import EDCore
from .gf import ImaginaryTimeGF, ImaginaryTimeGrid

__all__ = ["InteractionPair", "Expansion", "add_corr_operators"]


class InteractionPair:
    """`InteractionPair(operator_f, operator_i, propagator)` (src/expansion.jl:78-85):
    operator_f sits at the later time (arc tail), operator_i at the earlier time (arc head)."""

    def __init__(self, operator_f, operator_i, propagator: ImaginaryTimeGF):
        self.operator_f = np.asarray(operator_f)
        self.operator_i = np.asarray(operator_i)
        self.propagator = propagator

    def __getitem__(self, idx):  # pair[1] = operator_i, pair[2] = operator_f (1-based, :87-89)
        return (self.operator_i, self.operator_f)[idx - 1]


class Expansion:
    """Strong-coupling pseudo-particle expansion problem (src/expansion.jl:110-199)."""

    def __init__(self, ed: EDCore, grid: ImaginaryTimeGrid, interaction_pairs, corr_operators=(),
                 interpolate_ppgf=False):
        self.ed = ed
        self.grid = grid
        self.pairs = list(interaction_pairs)
        self.corr_operators = []
        # interpolate_ppgf only changes the *container* of P in the reference (IncSpline...); its
        # spline is never evaluated on this path, so it is accepted and ignored.
        self.interpolate_ppgf = bool(interpolate_ppgf)
        self.dims = list(ed.dims)
        self.S = len(self.dims)
        self.boff = np.concatenate([[0], np.cumsum([d * d for d in self.dims])]).astype(int)
        self.bsize = int(self.boff[-1])
        beta = grid.beta
        # P0: exact atomic propagator, lambda0 = log(Z_at)/beta (src/exact_atomic_ppgf.jl:130-135)
        self.lambda0 = np.log(ed.partition_function(beta)) / beta
        self.E_plus_lambda = [np.asarray(w) + self.lambda0 for w in ed.energies]
        # P: atomic PPGF on the grid, -i exp(-tau (E + lambda0)) (src/ppgf.jl:155-210)
        self.P = np.zeros((grid.n_tau, self.bsize), dtype=complex)
        for s, d in enumerate(self.dims):
            for i in range(d):
                self.P[:, self.boff[s] + i + d * i] = -1j * np.exp(-grid.tau * self.E_plus_lambda[s][i])
        # operator sector-block matrices (:168-178)
        self.identity_mat = ed.sector_block_matrix(ed.fock.identity())
        self.pair_operator_mat = [(ed.sector_block_matrix(p.operator_i),
                                   ed.sector_block_matrix(p.operator_f)) for p in self.pairs]
        self.corr_operators_mat = []
        # :180-183
        self.subspace_attachable_pairs = [
            [k for k, (op_i, _) in enumerate(self.pair_operator_mat) if s in op_i]
            for s in range(self.S)]
        for ops in corr_operators:
            add_corr_operators(self, ops)

    # -- block helpers -----------------------------------------------------------------------
    def block(self, packed, s):
        """Sector block s of a packed vector (-> [d, d]) or of a stack of them (-> [n, d, d])."""
        d = self.dims[s]
        a = np.asarray(packed)[..., self.boff[s]:self.boff[s + 1]]
        # column-major inside a block: element (i, j) sits at i + d*j
        return a.reshape(a.shape[:-1] + (d, d)).swapaxes(-1, -2)

    def P_sector(self, s):
        """P_s(tau_k) as an array [n_tau, d, d]."""
        return self.block(self.P, s)

    def pack_blocks(self, blocks):
        out = np.zeros(self.bsize, dtype=complex)
        for s, b in enumerate(blocks):
            out[self.boff[s]:self.boff[s + 1]] = np.asarray(b, dtype=complex).reshape(-1, order="F")
        return out

    # -- C-ABI payload -----------------------------------------------------------------------
    def flatten(self):
        """Arrays for `qiw_set_model` / `qiw_set_delta` (include/qinchworm.h)."""
        S = self.S
        sbms = []
        for op_i, op_f in self.pair_operator_mat:
            sbms += [op_i, op_f]
        n_pair_ops = len(sbms)
        for A, B in self.corr_operators_mat:
            sbms += [A, B]
        n_ops = len(sbms)
        op_target = -np.ones((n_ops, S), dtype=np.int32)
        op_mat_off = np.zeros((n_ops, S), dtype=np.int64)
        pool = []
        off = 0
        for o, sbm in enumerate(sbms):
            for s_i, (s_f, m) in sbm.items():
                op_target[o, s_i] = s_f
                op_mat_off[o, s_i] = off
                v = np.asarray(m, dtype=complex).reshape(-1, order="F")
                pool.append(v)
                off += v.size
        pool = np.concatenate(pool) if pool else np.zeros(0, dtype=complex)
        # distinct Delta tables: pairs sharing the same (kind, data) share a table
        tables, pair_table = [], []
        for p in self.pairs:
            g = p.propagator
            for t, (kind, data) in enumerate(tables):
                if kind == g.kind and data.shape == g.data.shape and np.array_equal(data, g.data):
                    pair_table.append(t)
                    break
            else:
                tables.append((g.kind, g.data))
                pair_table.append(len(tables) - 1)
        n_pairs = len(self.pairs)
        return dict(
            S=S, dims=np.asarray(self.dims, dtype=np.int32),
            energies=np.concatenate(self.E_plus_lambda).astype(np.float64),
            n_ops=n_ops, op_target=op_target, op_mat_off=op_mat_off,
            op_pool=np.ascontiguousarray(pool),
            n_pairs=n_pairs,
            pair_op_i=np.arange(0, 2 * n_pairs, 2, dtype=np.int32),
            pair_op_f=np.arange(1, 2 * n_pairs, 2, dtype=np.int32),
            pair_table=np.asarray(pair_table, dtype=np.int32),
            n_corr=len(self.corr_operators_mat),
            corr_A=np.arange(n_pair_ops, n_ops, 2, dtype=np.int32),
            corr_B=np.arange(n_pair_ops + 1, n_ops, 2, dtype=np.int32),
            tables=tables, beta=self.grid.beta, n_tau=self.grid.n_tau)


def add_corr_operators(expansion: Expansion, ops):
    """add_corr_operators!(expansion, (A, B)) — src/expansion.jl:302-310."""
    A, B = ops
    expansion.corr_operators.append((np.asarray(A), np.asarray(B)))
    expansion.corr_operators_mat.append((expansion.ed.sector_block_matrix(A),
                                         expansion.ed.sector_block_matrix(B)))
