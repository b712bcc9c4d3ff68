"""qMC inchworm drivers — host mirror of src/inchworm.jl on top of libqinchworm_cuda.so.

Same call shapes as the reference:
    inchworm_step_bare(expansion, grid, k_i, k_f, top_data)          src/inchworm.jl:228
    inchworm_step(expansion, grid, k_i, k_w, k_f, top_data)          src/inchworm.jl:123
    inchworm(expansion, grid, orders, orders_bare, N_samples, ...)   src/inchworm.jl:332
    correlator_2p(expansion, grid, orders, N_samples, ...)           src/inchworm.jl:918,1075
Grid points are passed as 0-based indices into `grid.tau`.  All sampling, interpolation and
diagram evaluation happens on the GPU inside `Context.eval`; what stays here is what stays in
Julia: topology tables, the sequential loop over time steps, set_ppgf!/normalize!, the
randomisation loop (one library call per scrambled sequence)."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import lib, ppgf
from .lib import MODE_BARE, MODE_BOLD, MODE_CORR

__all__ = ["RandomizationParams", "TopologiesInputData", "Solver", "inchworm_step", "inchworm_step_bare",
           "inchworm", "correlator_2p", "get_topologies_at_order"]


def get_topologies_at_order(order, k=None, with_external_arc=False):
    """src/diagrammatics.jl:322-337 (enumerated inside the library, same order and parity)."""
    return lib.topologies(order, k, with_external_arc)


@dataclass
class RandomizationParams:
    """src/randomization.jl:46-57.  rng: numpy Generator used to draw scrambling bits, or None to
    disable scrambling (the default, as in the reference)."""
    rng: object = None
    N_seqs: int = 1
    target_std: float = 0.0


@dataclass
class TopologiesInputData:
    """src/inchworm.jl:60-98."""
    order: int
    n_pts_after: int
    topologies: tuple            # (pairs[n, order, 2], parity[n])
    N_samples: int
    rand_params: RandomizationParams = field(default_factory=RandomizationParams)
    entry_id: int = -1           # handle of the compiled entry inside the library


def _scrambled_sequence(D, rng):
    """ScrambledSobolSeq(D, scramble_rng=rng) (src/scrambled_sobol.jl:66-143): the random bits are
    drawn here, in the order the reference draws them (shift first, then the matrices)."""
    m = lib.sobol_direction_numbers(D)
    if rng is None or D == 0:
        return m, np.zeros(D, dtype=np.uint32)
    shift_bits = rng.integers(0, 2, size=(D, 32), dtype=np.uint8)
    ltm_bits = rng.integers(0, 2, size=(D, 32, 32), dtype=np.uint8)
    return lib.sobol_scramble(m, shift_bits, ltm_bits)


class Solver:
    """Binds one `Expansion` to one library context (one GPU) and keeps compiled entries."""

    def __init__(self, expansion, ctx=None, device=-1):
        self.expansion = expansion
        self.ctx = ctx if ctx is not None else lib.Context(device=device)
        self.payload = self.ctx.set_expansion(expansion)
        self._next_entry = 0
        self._entries = {}       # (mode, order, n_pts_after, corr_idx) -> compiled entry, reused across calls
        self._steps = {}         # entry ids of a step -> _PreparedStep (host-stepped loop, default RandomizationParams)
        self._n_corr_uploaded = len(expansion.corr_operators_mat)

    def refresh_model(self):
        """Re-upload after add_corr_operators (invalidates compiled entries, as in the ABI)."""
        self.payload = self.ctx.set_expansion(self.expansion)
        self._next_entry = 0
        self._entries = {}
        self._steps = {}
        self._n_corr_uploaded = len(self.expansion.corr_operators_mat)

    def upload_P(self, first=0, count=None):
        count = self.expansion.grid.n_tau - first if count is None else count
        self.ctx.set_P(first, self.expansion.P[first:first + count])

    def make_entry(self, mode, order, n_pts_after, N_samples, rand_params=None, corr_idx=0):
        """TopologiesInputData for one (order, n_pts_after): the topology list is generated and compiled
        against the model once per Solver (the reference rebuilds both on every call,
        src/inchworm.jl:380,431-447 and :165)."""
        key = (mode, order, n_pts_after, corr_idx)
        if key in self._entries:
            eid, tops = self._entries[key]
            if eid is None:
                return None
            td = TopologiesInputData(order, n_pts_after, tops, N_samples, rand_params or RandomizationParams())
            td.entry_id = eid
            return td
        if mode == MODE_BARE:
            tops = get_topologies_at_order(order)
        else:
            tops = get_topologies_at_order(order, n_pts_after if order > 0 else 0, mode == MODE_CORR) \
                if order > 0 else get_topologies_at_order(0, 0, mode == MODE_CORR)
        if len(tops[1]) == 0:
            self._entries[key] = (None, tops)
            return None
        td = TopologiesInputData(order, n_pts_after, tops, N_samples, rand_params or RandomizationParams())
        td.entry_id = self._next_entry
        self._next_entry += 1
        self.ctx.set_topologies(td.entry_id, mode, order, n_pts_after, tops[0], tops[1], corr_idx=corr_idx)
        self._entries[key] = (td.entry_id, tops)
        return td

    def scale_P(self, k_f, value, lam=0.0):
        """The step seam (qiw_scale_P): row k_f of the device's table := value, then every row k times exp(-lam tau_k)."""
        if getattr(self, "_row", None) is None or self._row.shape[1] != self.ctx.bsize:
            self._row = np.zeros((1, self.ctx.bsize), dtype=np.complex128)
            self._row_ptr = lib._ptr(self._row.view(np.float64), lib.f64p)
        self._row[0] = value
        self.ctx.scale_P_prepared(k_f, self._row_ptr, float(lam))

    def eval_samples(self, t_i, t_w, t_f, top_data):
        """The randomised estimates of every entry's qMC integral: a list, per entry, of arrays [n_seqs_used, bsize]
        (mean_std_from_randomization is called per entry, src/inchworm.jl:142,174 and src/randomization.jl:86-100:
        every entry has its own early stop, and a host RNG stream is consumed entry by entry, N_seqs sequences each;
        order-0 entries are exact and draw nothing, src/inchworm.jl:148-157)."""
        if not top_data:
            return []
        ids = [td.entry_id for td in top_data]
        N = top_data[0].N_samples
        rp = top_data[0].rand_params
        assert all(td.N_samples == N for td in top_data)
        if rp.rng is None and rp.N_seqs == 1:
            # the default: one unscrambled sequence, every entry in ONE launch
            res = self.ctx.eval(t_i, t_w, t_f, ids, N)
            return [res[j][None, :] for j in range(len(ids))]
        if rp.target_std == 0.0:
            # no early stop (std == 0 exactly never triggers it in practice): all sequences of all entries in ONE launch;
            # the scrambling bits are drawn in the reference's order
            per_entry = [[_scrambled_sequence(2 * td.order, rp.rng if td.order > 0 else None) for _ in range(rp.N_seqs)]
                         for td in top_data]
            seqs = [[per_entry[j][s] for j in range(len(top_data))] for s in range(rp.N_seqs)]
            res = self.ctx.eval_seqs(t_i, t_w, t_f, ids, N, seqs)      # [n_seqs, n_entries, bsize]
            return [res[:1, j] if td.order == 0 else res[:, j] for j, td in enumerate(top_data)]
        out = []
        for td in top_data:          # early stop: entry by entry, one library call per sequence
            if td.order == 0:
                out.append(self.ctx.eval(t_i, t_w, t_f, [td.entry_id], N))
                continue
            smp = []
            for s_ in range(rp.N_seqs):
                seq = _scrambled_sequence(2 * td.order, rp.rng)
                smp.append(self.ctx.eval(t_i, t_w, t_f, [td.entry_id], N, sobol=[seq])[0])
                if s_ > 0 and np.max(np.abs(np.std(smp, axis=0, ddof=1))) <= rp.target_std:
                    break
            out.append(np.array(smp))
        return out

    def eval_entries(self, t_i, t_w, t_f, top_data):
        """mean/std over randomised sequences of the per-entry qMC integrals
        (mean_std_from_randomization, src/randomization.jl:86-100).  Returns (mean, std) arrays
        [n_entries, bsize]; std is NaN for a single sequence, as in the reference, and 0 for the exact order-0 entries."""
        if not top_data:
            z = np.zeros((0, self.ctx.bsize), dtype=complex)
            return z, z
        rp = top_data[0].rand_params
        if rp.rng is None and rp.N_seqs == 1:
            # the default (one unscrambled sequence): the estimate itself, no statistics — one library call, no per-entry work
            mean = self.ctx.eval(t_i, t_w, t_f, [td.entry_id for td in top_data], top_data[0].N_samples)
            std = np.full_like(mean, np.nan)
            std[[j for j, td in enumerate(top_data) if td.order == 0]] = 0.0  # exact evaluation (src/inchworm.jl:155)
            return mean, std
        samples = self.eval_samples(t_i, t_w, t_f, top_data)
        mean = np.array([x.mean(axis=0) for x in samples])
        std = np.full_like(mean, np.nan)
        for j, (td, x) in enumerate(zip(top_data, samples)):
            if td.order == 0:
                std[j] = 0.0  # exact evaluation (src/inchworm.jl:155)
            elif len(x) > 1:
                std[j] = np.std(x, axis=0, ddof=1)
        return mean, std


def _order_sums(top_data, mean, std, bsize):
    orders = sorted({td.order for td in top_data})
    ord_of = np.array([td.order for td in top_data])
    contribs = {o: mean[ord_of == o].sum(axis=0) if len(mean) else np.zeros(bsize, dtype=complex) for o in orders}
    contribs_std = {o: std[ord_of == o].sum(axis=0) if len(std) else np.zeros(bsize, dtype=complex) for o in orders}
    total = sum(contribs.values()) if contribs else np.zeros(bsize, dtype=complex)
    return total, contribs, contribs_std


class _PreparedStep:
    """What one host-stepped step with the default RandomizationParams needs besides the library call, built once per
    list of entries: the id array, the result buffer and their ctypes pointers, and the per-order sums as ONE real
    matmul onehot[order, entry] x result[entry, (element, re/im)] (the interpreter, not the GPU, bounds this seam)."""

    def __init__(self, ctx, top_data):
        self.n = len(top_data)
        self.ids = np.ascontiguousarray([td.entry_id for td in top_data], dtype=np.int32)
        self.out = np.zeros((self.n, ctx.bsize), dtype=np.complex128)
        self.out_real = self.out.view(np.float64)
        self.ids_ptr, self.out_ptr = lib._ptr(self.ids, lib.i32p), lib._ptr(self.out_real, lib.f64p)
        self.orders = sorted({td.order for td in top_data})
        self.onehot = np.zeros((len(self.orders), self.n))
        for j, td in enumerate(top_data):
            self.onehot[self.orders.index(td.order), j] = 1.0
        # the std of a single sequence is NaN (src/randomization.jl:99), of the exact order-0 entries 0 (src/inchworm.jl:155)
        self.std = {o: np.full(ctx.bsize, 0.0 if o == 0 else np.nan, dtype=complex) for o in self.orders}


def _step(solver: Solver, t_i, t_w, t_f, top_data):
    rp = top_data[0].rand_params if top_data else None
    if rp is not None and rp.rng is None and rp.N_seqs == 1:
        key = tuple(td.entry_id for td in top_data)
        ps = solver._steps.get(key)
        if ps is None:
            ps = solver._steps[key] = _PreparedStep(solver.ctx, top_data)
        N = top_data[0].N_samples       # (not cached: a Solver reuses compiled entries across calls with other sample counts)
        assert all(td.N_samples == N for td in top_data)
        solver.ctx.eval_prepared(t_i, t_w, t_f, ps.n, ps.ids_ptr, N, ps.out_ptr)
        sums = (ps.onehot @ ps.out_real).view(np.complex128)          # [n_orders, bsize], a fresh array
        return sums.sum(axis=0), {o: sums[q] for q, o in enumerate(ps.orders)}, {o: v.copy() for o, v in ps.std.items()}
    mean, std = solver.eval_entries(t_i, t_w, t_f, top_data)
    return _order_sums(top_data, mean, std, solver.ctx.bsize)


def inchworm_step_bare(solver: Solver, grid, k_i, k_f, top_data):
    """One initial step with bare propagators (src/inchworm.jl:228-304).  Returns
    (value, order_contribs, order_contribs_std) as packed block vectors."""
    return _step(solver, grid.tau[k_i], grid.tau[k_i], grid.tau[k_f], top_data)


def inchworm_step(solver: Solver, grid, k_i, k_w, k_f, top_data):
    """One regular inchworm step with bold propagators (src/inchworm.jl:123-204)."""
    return _step(solver, grid.tau[k_i], grid.tau[k_w], grid.tau[k_f], top_data)


def _bold_entries(solver, orders, N_samples, rand_params, n_pts_after_max):
    top_data = []
    for o in orders:
        rng = [0] if o == 0 else range(1, min(2 * o - 1, n_pts_after_max or 10 ** 9) + 1)
        for k in rng:
            td = solver.make_entry(MODE_BOLD, o, k, N_samples, rand_params)
            if td is not None:
                top_data.append(td)
    return top_data


def inchworm(expansion, grid, orders, orders_bare, N_samples, n_pts_after_max=None,
             rand_params=None, solver=None, device_resident=None):
    """inchworm!(expansion, grid, orders, orders_bare, N_samples; ...) (src/inchworm.jl:332-498).
    Results are written into `expansion.P`; returns (P_orders, P_orders_std): dicts
    order -> [n_tau, bsize] arrays of order-resolved contributions.

    device_resident=True runs the whole loop inside the library (qiw_inchworm_run): set_ppgf! and
    normalize! happen on the GPU between steps, with no host round trip.  It requires the default
    RandomizationParams (one unscrambled sequence reused at every step) and is chosen automatically
    in that case (device_resident=None); device_resident=False forces the host-stepped loop, one
    qiw_eval per step, which is what the Julia shim does."""
    assert N_samples == 0 or (N_samples & (N_samples - 1)) == 0, "N_samples must be a power of 2"
    rand_params = rand_params or RandomizationParams()
    assert rand_params.N_seqs > 0
    solver = solver or Solver(expansion)
    if device_resident is None:
        device_resident = rand_params.rng is None and rand_params.N_seqs == 1
    n_tau = grid.n_tau
    orders, orders_bare = list(orders), list(orders_bare)
    if N_samples == 0:
        # the reference skips every sampled entry (`td.N_samples <= 0 && continue`, src/inchworm.jl:159,258) and
        # evaluates order 0 only; the exact entries need no samples, the library wants a positive count
        orders, orders_bare = [o for o in orders if o == 0], [o for o in orders_bare if o == 0]
        N_samples = 1
    P_orders = {o: np.zeros((n_tau, solver.ctx.bsize), dtype=complex) for o in set(orders) | set(orders_bare)}
    P_orders_std = {o: np.zeros((n_tau, solver.ctx.bsize), dtype=complex) for o in P_orders}
    if device_resident:
        assert rand_params.rng is None and rand_params.N_seqs == 1, "device-resident loop: default RandomizationParams only"
        bare = [solver.make_entry(MODE_BARE, o, 2 * o, N_samples, rand_params) for o in orders_bare]
        bold = _bold_entries(solver, orders, N_samples, rand_params, n_pts_after_max)
        solver.upload_P()
        hist = solver.ctx.inchworm_run([td.entry_id for td in bare], [td.entry_id for td in bold], N_samples)
        expansion.P[:] = solver.ctx.get_P()
        # per-order sums of the per-entry history: one real batched matmul, onehot[order, entry] x hist[k, entry, (el, re/im)]
        o_list = sorted(P_orders)
        onehot = np.zeros((len(o_list), hist.shape[1]))
        for j, td in enumerate(bare + bold):
            onehot[o_list.index(td.order), j] = 1.0
        hr = np.ascontiguousarray(hist).view(np.float64).reshape(hist.shape[0], hist.shape[1], 2 * hist.shape[2])
        sums = np.matmul(onehot, hr).reshape(hist.shape[0], len(o_list), hist.shape[2], 2).view(np.complex128)[..., 0]
        for q, o in enumerate(o_list):
            P_orders[o] += sums[:, q, :]
        for o in P_orders_std:         # std of a single sequence is NaN (src/randomization.jl:99); order 0 is exact
            if o > 0:                  # (:155), and grid point 0 is never evaluated
                P_orders_std[o][1:] = np.nan
        return P_orders, P_orders_std
    # first step: bare diagrams (:373-416)
    top_data = [solver.make_entry(MODE_BARE, o, 2 * o, N_samples, rand_params) for o in orders_bare]
    solver.upload_P()
    value, contribs, contribs_std = inchworm_step_bare(solver, grid, 0, 1, top_data)
    ppgf.set_ppgf(expansion, 1, value)
    for o in contribs:
        P_orders[o][1] = contribs[o]
        P_orders_std[o][1] = contribs_std[o]
    # the rest of inching (:420-493)
    top_data = _bold_entries(solver, orders, N_samples, rand_params, n_pts_after_max)
    solver.scale_P(1, value)
    for n in range(1, n_tau - 1):
        value, contribs, contribs_std = inchworm_step(solver, grid, 0, n, n + 1, top_data)
        ppgf.set_ppgf(expansion, n + 1, value)
        lam = ppgf.normalize_at(expansion, n + 1)  # suppress exponential growth (:488)
        # the device's table follows with one row and lambda (qiw_scale_P) instead of a re-upload of the whole table
        solver.scale_P(n + 1, value, lam)
        for o in contribs:
            P_orders[o][n + 1] = contribs[o]
            P_orders_std[o][n + 1] = contribs_std[o]
    return P_orders, P_orders_std


def correlator_2p(expansion, grid, orders, N_samples, rand_params=None, solver=None, return_std=False, batch=True):
    """correlator_2p(expansion, grid, orders, N_samples) (src/inchworm.jl:918-1086): one array
    [n_tau] per registered pair in expansion.corr_operators.  With the default RandomizationParams all grid
    points of a pair are evaluated by one qiw_eval_batch launch (batch=False: one qiw_eval per point)."""
    assert N_samples == 0 or (N_samples & (N_samples - 1)) == 0
    rand_params = rand_params or RandomizationParams()
    solver = solver or Solver(expansion)
    if N_samples == 0:       # sampled entries are skipped (src/inchworm.jl:843): order 0 only
        orders, N_samples = [o for o in orders if o == 0], 1
    if solver._n_corr_uploaded != len(expansion.corr_operators_mat):
        solver.refresh_model()
    solver.upload_P()
    n_tau = grid.n_tau
    diag = ppgf._diag_indices(expansion)
    Z = ppgf.partition_function(expansion)
    out, out_std = [], []
    for c in range(len(expansion.corr_operators)):
        top_data = []
        for o in orders:
            for k in ([0] if o == 0 else range(1, 2 * o)):
                td = solver.make_entry(MODE_CORR, o, k, N_samples, rand_params, corr_idx=c)
                if td is not None:
                    top_data.append(td)
        g = np.zeros(n_tau, dtype=complex)
        g_std = np.zeros(n_tau, dtype=complex)
        batched = batch and rand_params.rng is None and rand_params.N_seqs == 1 and n_tau > 1 and top_data
        for k in range(n_tau):
            if batched and k > 0:
                break
            tds = top_data
            if k == 0:  # only order 0 contributes at tau_A = tau_B (:1013-1024)
                tds = top_data[:1] if top_data and top_data[0].order == 0 else []
                if not tds:
                    continue
            samples = solver.eval_samples(grid.tau[0], grid.tau[k], grid.tau[-1], tds)
            # per entry: mean and std over the sequences of tr(...) (the reference integrates the trace, :862-872);
            # sums over the entries, divided by the partition function (:882-889)
            for td, x in zip(tds, samples):
                trs = x[:, diag].sum(axis=1)
                g[k] += trs.mean() / Z
                g_std[k] += (0.0 if td.order == 0 else (np.std(trs, ddof=1) if len(trs) > 1 else np.nan)) / Z
        if batched:
            # every grid point tau_k, k >= 1, in ONE launch: the (pair, tau) evaluations are independent (:995-1046)
            times = np.array([[grid.tau[0], grid.tau[k], grid.tau[-1]] for k in range(1, n_tau)])
            res = solver.ctx.eval_batch(times, [td.entry_id for td in top_data], N_samples)
            g[1:] = res[:, :, diag].sum(axis=(1, 2)) / Z
            g_std[1:] = np.nan if any(td.order > 0 for td in top_data) else 0.0   # std of one sequence (src/randomization.jl:99)
        out.append(g)
        out_std.append(g_std)
    return (out, out_std) if return_std else out
