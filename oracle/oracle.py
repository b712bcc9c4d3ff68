"""ctypes harness around oracle/libqiw_oracle.so plus reference-shaped drivers — TEST
INFRASTRUCTURE ONLY (tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference).

The drivers restate the reference's host loops on top of the C++ oracle so that the whole
CPU path is independent of the product's host layer:
  inchworm_step_bare : src/inchworm.jl:228-304      inchworm     : src/inchworm.jl:332-498
  inchworm_step      : src/inchworm.jl:123-204      correlator_2p: src/inchworm.jl:805-890,918-1052
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libqiw_oracle.so")

MODE_BARE, MODE_BOLD, MODE_CORR = 0, 1, 2


def build(force=False):
    src = os.path.join(HERE, "qiw_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"] + (["-B"] if force else []))
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.qo_create.restype = C.c_void_p
        L.qo_last_error.restype = C.c_char_p
        L.qo_set_and_normalize.restype = C.c_double
        for name in ("qo_destroy", "qo_last_error", "qo_set_threads", "qo_set_pool", "qo_set_model", "qo_set_delta",
                     "qo_set_grid", "qo_set_P", "qo_get_P", "qo_set_topologies", "qo_eval",
                     "qo_eval_at_times", "qo_last_counts", "qo_set_and_normalize"):
            getattr(L, name).argtypes = None
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _c2d(a):
    """complex array -> contiguous float64 view (re, im interleaved)."""
    a = np.ascontiguousarray(a, dtype=np.complex128)
    return a, a.view(np.float64)


# ---------------------------------------------------------------------------------------------
# Sobol / topologies / transforms (stateless helpers)
# ---------------------------------------------------------------------------------------------

def sobol_direction_numbers(D):
    m = np.zeros((max(D, 1), 32), dtype=np.uint32)
    rc = lib().qo_sobol_direction_numbers(C.c_int(D), _p(m, C.c_uint32))
    if rc:
        raise ValueError("Invalid Sobol dimension %d" % D)
    return m[:D]


def sobol_scramble(m, shift_bits, ltm_bits):
    """LMS+shift from explicit random bits; shift_bits [D,32], ltm_bits [D,32,32] indexed as the
    reference's Julia arrays (column-major fill happens in the caller)."""
    D = m.shape[0]
    m = np.ascontiguousarray(m, dtype=np.uint32).copy()
    x0 = np.zeros(D, dtype=np.uint32)
    sb = np.asfortranarray(shift_bits, dtype=np.uint8)
    lb = np.asfortranarray(ltm_bits, dtype=np.uint8)
    lib().qo_sobol_scramble(C.c_int(D), _p(m, C.c_uint32), _p(x0, C.c_uint32),
                            sb.ctypes.data_as(C.POINTER(C.c_uint8)), lb.ctypes.data_as(C.POINTER(C.c_uint8)))
    return m, x0


def sobol_points(m, x0, skip, count):
    D = m.shape[0]
    xi = np.zeros((count, D), dtype=np.uint32)
    xf = np.zeros((count, D), dtype=np.float64)
    m = np.ascontiguousarray(m, dtype=np.uint32)
    x0 = np.ascontiguousarray(x0, dtype=np.uint32)
    lib().qo_sobol_points(C.c_int(D), _p(m, C.c_uint32), _p(x0, C.c_uint32), C.c_uint64(skip),
                          C.c_uint64(count), _p(xi, C.c_uint32), _p(xf, C.c_double))
    return xi, xf


def topologies(order, k=None, with_external_arc=False):
    kk = -1 if k is None else int(k)
    n = lib().qo_topologies(C.c_int(order), C.c_int(kk), C.c_int(int(with_external_arc)), None, None)
    pairs = np.zeros((n, order, 2), dtype=np.int32)
    parity = np.zeros(n, dtype=np.int32)
    if n:
        lib().qo_topologies(C.c_int(order), C.c_int(kk), C.c_int(int(with_external_arc)),
                            _p(pairs, C.c_int32), _p(parity, C.c_int32))
    return pairs, parity


def transform(mode, d_before, d_after, t_i, t_w, t_f, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    u = np.zeros_like(x)
    jac = C.c_double(0)
    lib().qo_transform(C.c_int(mode), C.c_int(d_before), C.c_int(d_after), C.c_double(t_i),
                       C.c_double(t_w), C.c_double(t_f), _p(x, C.c_double), _p(u, C.c_double), C.byref(jac))
    return u, jac.value


def rank_sub_range(N, n_ranks, rank):
    s, c = C.c_uint64(0), C.c_uint64(0)
    lib().qo_rank_sub_range(C.c_uint64(N), C.c_int(n_ranks), C.c_int(rank), C.byref(s), C.byref(c))
    return s.value, c.value


# ---------------------------------------------------------------------------------------------
# Stateful oracle bound to one flattened Expansion payload
# ---------------------------------------------------------------------------------------------

class Oracle:
    def __init__(self, payload, P, threads=1):
        """payload: dict as produced by Expansion.flatten(); P: [n_tau, bsize] complex."""
        L = lib()
        self.L = L
        self.h = C.c_void_p(L.qo_create())
        self.bsize = int(sum(int(d) ** 2 for d in payload["dims"]))
        self.n_tau = int(payload["n_tau"])
        self.beta = float(payload["beta"])
        self.entries = {}
        L.qo_set_threads(self.h, C.c_int(threads))
        pl = payload
        pool, poolv = _c2d(pl["op_pool"])
        dims = np.ascontiguousarray(pl["dims"], dtype=np.int32)
        en = np.ascontiguousarray(pl["energies"], dtype=np.float64)
        tgt = np.ascontiguousarray(pl["op_target"], dtype=np.int32)
        off = np.ascontiguousarray(pl["op_mat_off"], dtype=np.int64)
        pi = np.ascontiguousarray(pl["pair_op_i"], dtype=np.int32)
        pf = np.ascontiguousarray(pl["pair_op_f"], dtype=np.int32)
        pt = np.ascontiguousarray(pl["pair_table"], dtype=np.int32)
        ca = np.ascontiguousarray(pl["corr_A"], dtype=np.int32)
        cb = np.ascontiguousarray(pl["corr_B"], dtype=np.int32)
        L.qo_set_model(self.h, C.c_int(pl["S"]), _p(dims, C.c_int32), _p(en, C.c_double),
                       C.c_int(pl["n_ops"]), _p(tgt, C.c_int32), _p(off, C.c_int64),
                       poolv.ctypes.data_as(C.POINTER(C.c_double)), C.c_int(pl["n_pairs"]),
                       _p(pi, C.c_int32), _p(pf, C.c_int32), _p(pt, C.c_int32),
                       C.c_int(pl["n_corr"]), _p(ca, C.c_int32), _p(cb, C.c_int32))
        for t, (kind, data) in enumerate(pl["tables"]):
            d, dv = _c2d(data)
            L.qo_set_delta(self.h, C.c_int(t), C.c_int(kind), C.c_int(len(d)), C.c_double(self.beta),
                           dv.ctypes.data_as(C.POINTER(C.c_double)))
        L.qo_set_grid(self.h, C.c_int(self.n_tau), C.c_double(self.beta))
        self.set_P(0, P)

    def __del__(self):
        try:
            self.L.qo_destroy(self.h)
        except Exception:
            pass

    def set_threads(self, n):
        self.L.qo_set_threads(self.h, C.c_int(n))

    def set_pool(self, on):
        """True (default): persistent worker threads with cached evaluators; False: a thread per (entry, call)."""
        self.L.qo_set_pool(self.h, C.c_int(int(on)))

    def set_P(self, first, rows):
        r, rv = _c2d(np.atleast_2d(rows))
        self.L.qo_set_P(self.h, C.c_int(first), C.c_int(r.shape[0]), rv.ctypes.data_as(C.POINTER(C.c_double)))

    def get_P(self):
        out = np.zeros((self.n_tau, self.bsize), dtype=np.complex128)
        self.L.qo_get_P(self.h, C.c_int(0), C.c_int(self.n_tau), out.view(np.float64).ctypes.data_as(C.POINTER(C.c_double)))
        return out

    def set_topologies(self, entry_id, mode, order, n_pts_after, pairs, parity):
        pairs = np.ascontiguousarray(pairs, dtype=np.int32)
        parity = np.ascontiguousarray(parity, dtype=np.int32)
        self.entries[entry_id] = (mode, order, n_pts_after, len(parity))
        self.L.qo_set_topologies(self.h, C.c_int(entry_id), C.c_int(mode), C.c_int(order),
                                 C.c_int(n_pts_after), C.c_int(len(parity)),
                                 _p(pairs, C.c_int32), _p(parity, C.c_int32))

    def eval(self, t_i, t_w, t_f, entry_ids, N_total, start=0, count=None, corr_idx=0, sobol=None):
        """Returns [n_entries, bsize] complex: per-entry qMC integrals restricted to the Sobol
        index range [start, start+count), normalised by N_total."""
        count = N_total - start if count is None else count
        ids = np.ascontiguousarray(entry_ids, dtype=np.int32)
        ms, xs = [], []
        for j, e in enumerate(ids):
            order = self.entries[int(e)][1]
            if sobol is not None:
                m, x0 = sobol[j]
            else:
                m = sobol_direction_numbers(2 * order)
                x0 = np.zeros(2 * order, dtype=np.uint32)
            ms.append(np.asarray(m, dtype=np.uint32).reshape(-1))
            xs.append(np.asarray(x0, dtype=np.uint32).reshape(-1))
        mcat = np.ascontiguousarray(np.concatenate(ms + [np.zeros(1, np.uint32)]))
        xcat = np.ascontiguousarray(np.concatenate(xs + [np.zeros(1, np.uint32)]))
        out = np.zeros((len(ids), self.bsize), dtype=np.complex128)
        rc = self.L.qo_eval(self.h, C.c_double(t_i), C.c_double(t_w), C.c_double(t_f), C.c_int(corr_idx),
                            C.c_int(len(ids)), _p(ids, C.c_int32), _p(mcat, C.c_uint32), _p(xcat, C.c_uint32),
                            C.c_uint64(start), C.c_uint64(count), C.c_uint64(N_total),
                            out.view(np.float64).ctypes.data_as(C.POINTER(C.c_double)))
        if rc:
            raise RuntimeError(self.L.qo_last_error(self.h).decode())
        return out

    def eval_at_times(self, entry_id, t_i, t_w, t_f, times, corr_idx=0):
        times = np.ascontiguousarray(times, dtype=np.float64)
        out = np.zeros((times.shape[0], self.bsize), dtype=np.complex128)
        rc = self.L.qo_eval_at_times(self.h, C.c_int(entry_id), C.c_double(t_i), C.c_double(t_w), C.c_double(t_f),
                                     C.c_int(corr_idx), C.c_int(times.shape[0]), _p(times, C.c_double),
                                     out.view(np.float64).ctypes.data_as(C.POINTER(C.c_double)))
        if rc:
            raise RuntimeError("block off-diagonal contribution")
        return out

    def last_counts(self):
        f, l = C.c_double(0), C.c_double(0)
        self.L.qo_last_counts(self.h, C.byref(f), C.byref(l))
        return f.value, l.value

    def set_and_normalize(self, k_f, value, do_normalize=True):
        v, vv = _c2d(value)
        return self.L.qo_set_and_normalize(self.h, C.c_int(k_f), vv.ctypes.data_as(C.POINTER(C.c_double)),
                                           C.c_int(int(do_normalize)))


# ---------------------------------------------------------------------------------------------
# Reference-shaped drivers
# ---------------------------------------------------------------------------------------------

def _diag_idx(dims):
    idx, off = [], 0
    for d in dims:
        idx += [off + i + d * i for i in range(d)]
        off += d * d
    return np.asarray(idx)


def inchworm(payload, P0_table, orders, orders_bare, N_samples, n_pts_after_max=None, threads=1,
             n_ranks=1, max_bold_steps=None, pool=True):
    """inchworm!(expansion, grid, orders, orders_bare, N_samples) on the oracle.

    Returns dict(P=[n_tau,bsize] final (per-step normalised) table, P_orders={order: [n_tau,bsize]},
    evals=diagram evaluations performed, oracle=Oracle).  n_ranks > 1 emulates the MPI split
    (src/mpi.jl:49-54): every "rank" evaluates its sub-range and the partial sums are added.
    """
    o = Oracle(payload, P0_table, threads=threads)
    o.set_pool(pool)
    n_tau, beta = o.n_tau, o.beta
    tau = np.linspace(0.0, beta, n_tau)
    orders, orders_bare = list(orders), list(orders_bare)
    P_orders = {k: np.zeros((n_tau, o.bsize), dtype=complex) for k in set(orders) | set(orders_bare)}
    evals = 0

    def run(ids, t_i, t_w, t_f):
        if n_ranks == 1:
            return o.eval(t_i, t_w, t_f, ids, N_samples)
        tot = 0
        for r in range(n_ranks):
            s, c = rank_sub_range(N_samples, n_ranks, r)
            part = o.eval(t_i, t_w, t_f, ids, N_samples, start=s, count=c)
            if r > 0:  # order-0 entries are exact and identical on every rank: count them once
                for j, e in enumerate(ids):
                    if o.entries[int(e)][1] == 0:
                        part[j] = 0
            tot = tot + part
        return tot

    # bare step (src/inchworm.jl:373-416)
    eid = 0
    bare_ids = []
    for order in orders_bare:
        pairs, parity = topologies(order)
        o.set_topologies(eid, MODE_BARE, order, 2 * order, pairs, parity)
        bare_ids.append(eid); eid += 1
        evals += (N_samples if order > 0 else 1) * len(parity)
    res = run(bare_ids, tau[0], tau[0], tau[1])
    for j, order in enumerate(orders_bare):
        P_orders[order][1] += res[j]
    o.set_and_normalize(1, res.sum(axis=0), do_normalize=False)
    # bold steps (:420-493)
    bold_ids, bold_orders, n_top_bold = [], [], 0
    for order in orders:
        rng = [0] if order == 0 else range(1, min(2 * order - 1, n_pts_after_max or 10 ** 9) + 1)
        for k in rng:
            pairs, parity = topologies(order, k)
            if len(parity) == 0:
                continue
            o.set_topologies(eid, MODE_BOLD, order, k, pairs, parity)
            bold_ids.append(eid); bold_orders.append(order); eid += 1
            n_top_bold += (N_samples if order > 0 else 1) * len(parity)
    for n in range(1, n_tau - 1):  # Julia n = 2 : n_tau-1  ->  tau_w = tau[n], tau_f = tau[n+1]
        if max_bold_steps is not None and n > max_bold_steps:
            break  # bounded sample for CPU-baseline timing
        res = run(bold_ids, tau[0], tau[n], tau[n + 1])
        for j, order in enumerate(bold_orders):
            P_orders[order][n + 1] += res[j]
        o.set_and_normalize(n + 1, res.sum(axis=0), do_normalize=True)
        evals += n_top_bold
    return dict(P=o.get_P(), P_orders=P_orders, evals=evals, oracle=o)


def correlator_2p(payload, P_table, orders, N_samples, corr_idx=0, threads=1, tau_indices=None):
    """correlator_2p(expansion, grid, orders, N_samples) for one registered (A, B) pair."""
    o = Oracle(payload, P_table, threads=threads)
    n_tau, beta = o.n_tau, o.beta
    tau = np.linspace(0.0, beta, n_tau)
    ids, ords = [], []
    eid = 0
    for order in orders:
        rng = [0] if order == 0 else range(1, 2 * order)
        for k in rng:
            pairs, parity = topologies(order, k, with_external_arc=True)
            if len(parity) == 0:
                continue
            o.set_topologies(eid, MODE_CORR, order, k, pairs, parity)
            ids.append(eid); ords.append(order); eid += 1
    dg = _diag_idx(payload["dims"])
    Z = (1j * np.asarray(P_table)[-1, dg]).sum()
    out = np.zeros(n_tau, dtype=complex)
    ks = range(n_tau) if tau_indices is None else tau_indices
    for k in ks:
        if k == 0:  # only order 0 contributes at tau_A = tau_B (src/inchworm.jl:1013-1024)
            if ords and ords[0] == 0:
                r = o.eval(tau[0], tau[0], tau[-1], ids[:1], N_samples, corr_idx=corr_idx)
                out[0] = r[:, dg].sum() / Z
            continue
        r = o.eval(tau[0], tau[k], tau[-1], ids, N_samples, corr_idx=corr_idx)
        out[k] = r[:, dg].sum() / Z
    return out
