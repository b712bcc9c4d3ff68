"""Tiny exact-diagonalisation helper standing in for KeldyshED.jl (host side, set-up only).

The reference obtains sectors, eigen-energies and operator blocks from `KeldyshED.EDCore`
(src/expansion.jl:147-183, src/sector_block_matrix.jl:52-66).  KeldyshED is a third-party Julia
package that is not in this image, so synthetic models for tests and benchmarks are built here:
dense operators on the Fock space of <= ~6 fermionic modes, invariant subspaces found by the same
two-phase auto-partition idea (connected components of H plus symmetry breakers, then merging so
that every c / c^dagger maps a subspace into exactly one subspace), eigenbasis per subspace.

Nothing here is on the hot path; it only produces the payload of `qiw_set_model`.
"""
from __future__ import annotations

import numpy as np

__all__ = ["FockSpace", "EDCore"]


class FockSpace:
    """Fock space of `labels` fermionic modes; bit i of a state index = occupation of mode i.

    Sign convention: c^dag_i |n> = (-1)^{sum_{j<i} n_j} |n + e_i>.
    """

    def __init__(self, labels):
        self.labels = [tuple(l) if isinstance(l, (list, tuple)) else (l,) for l in labels]
        self.n = len(self.labels)
        self.dim = 1 << self.n
        self._idx = {l: i for i, l in enumerate(self.labels)}

    def _mode(self, *label):
        return self._idx[tuple(label)]

    def c(self, *label):
        i = self._mode(*label)
        m = np.zeros((self.dim, self.dim))
        for s in range(self.dim):
            if s >> i & 1:
                sign = (-1) ** bin(s & ((1 << i) - 1)).count("1")
                m[s ^ (1 << i), s] = sign
        return m

    def c_dag(self, *label):
        return self.c(*label).T.copy()

    def n_op(self, *label):
        return self.c_dag(*label) @ self.c(*label)

    def identity(self):
        return np.eye(self.dim)


class EDCore:
    """Block-diagonalised atomic problem: `subspaces`, `energies`, `unitaries`.

    energies are shifted so that the ground state has E = 0 (KeldyshED convention; the shift is
    immaterial for P0 because lambda0 absorbs it, src/exact_atomic_ppgf.jl:130-135).
    """

    def __init__(self, fock: FockSpace, H, symmetry_breakers=()):
        self.fock = fock
        H = np.asarray(H)
        dim = fock.dim
        # Phase 1: connected components of |H| + sum |breakers|
        conn = np.abs(H) > 1e-14
        for b in symmetry_breakers:
            conn |= np.abs(np.asarray(b)) > 1e-14
        parent = list(range(dim))

        def find(a):
            while parent[a] != a:
                parent[a] = parent[parent[a]]
                a = parent[a]
            return a

        def union(a, b):
            ra, rb = find(a), find(b)
            if ra != rb:
                parent[max(ra, rb)] = min(ra, rb)

        for i, j in zip(*np.nonzero(conn)):
            union(int(i), int(j))
        # Phase 2: merge until every c / c^dag maps a subspace into exactly one subspace
        ops = [fock.c(*l) for l in fock.labels] + [fock.c_dag(*l) for l in fock.labels]
        changed = True
        while changed:
            changed = False
            for o in ops:
                image = {}  # root of source subspace -> one Fock state of its image
                for src in range(dim):
                    for t in np.nonzero(np.abs(o[:, src]) > 1e-14)[0]:
                        r = find(src)
                        if r in image:
                            if find(image[r]) != find(int(t)):
                                union(image[r], int(t))
                                changed = True
                        else:
                            image[r] = int(t)
                if changed:
                    break  # roots moved: rebuild the image map from scratch
        roots = sorted({find(i) for i in range(dim)})
        self.subspaces = [[i for i in range(dim) if find(i) == r] for r in roots]
        # eigen-decomposition per subspace
        self.energies, self.unitaries = [], []
        for sp in self.subspaces:
            h = H[np.ix_(sp, sp)]
            w, v = np.linalg.eigh((h + h.conj().T) / 2)
            self.energies.append(w)
            self.unitaries.append(v)
        gs = min(w.min() for w in self.energies)
        self.energies = [w - gs for w in self.energies]
        self.gs_energy = gs
        self.dims = [len(sp) for sp in self.subspaces]

    def partition_function(self, beta):
        return float(sum(np.exp(-beta * w).sum() for w in self.energies))

    def density_matrix(self, beta):
        z = self.partition_function(beta)
        return [np.diag(np.exp(-beta * w)) / z for w in self.energies]

    def operator_blocks(self, O):
        """Dict (s_f, s_i) -> block of O in the eigenbasis (KeldyshED.operator_blocks)."""
        O = np.asarray(O)
        out = {}
        for si, (spi, ui) in enumerate(zip(self.subspaces, self.unitaries)):
            for sf, (spf, uf) in enumerate(zip(self.subspaces, self.unitaries)):
                blk = uf.conj().T @ O[np.ix_(spf, spi)] @ ui
                if np.abs(blk).max() > 1e-13:
                    out[(sf, si)] = blk.astype(complex)
        return out

    def sector_block_matrix(self, O):
        """Dict s_i -> (s_f, block); raises if a column has more than one block
        (src/sector_block_matrix.jl:52-66)."""
        sbm = {}
        for (sf, si), m in self.operator_blocks(O).items():
            if si in sbm:
                raise ValueError("operator is not representable by a SectorBlockMatrix "
                                 "(more than one non-zero block per column)")
            sbm[si] = (sf, m)
        return sbm

    def to_fock_basis(self, blocks):
        """Block-diagonal list in the eigenbasis -> dense matrix in the Fock basis."""
        out = np.zeros((self.fock.dim, self.fock.dim), dtype=complex)
        for sp, u, b in zip(self.subspaces, self.unitaries, blocks):
            out[np.ix_(sp, sp)] = u @ np.asarray(b) @ u.conj().T
        return out
