"""CPU-side checks of the product's host logic: the C-ABI library loads and exports every symbol
include/qinchworm.h declares, the integer work (Sobol tables, scrambling, topology enumeration,
sample partitioning) is bit-exact against the oracle, and the host-side topology compiler
produces programs that replay to the oracle's values.  No GPU compute is called."""
import ctypes
import os
import re

import numpy as np
import pytest

import models
from program_interp import run_lane_program, run_program, run_records


def test_library_exports_every_declared_symbol(qlib):
    hdr = open(qlib.HEADER_PATH).read()
    declared = set(re.findall(r"\b(qiw_[A-Za-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 25
    L = ctypes.CDLL(qlib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), "symbol %s declared in qinchworm.h but not exported" % name
    assert declared == set(qlib.exported_symbols())
    assert qlib.load().qiw_version().decode() == "0.1.0"


def test_no_cpu_fallback(qlib):
    """Without a device the compute path must fail loudly; the planning-only context must refuse
    every compute entry point."""
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(qlib.QiwError):
            qlib.Context()
    ex, grid, f = models.single_level(n_tau=10, spline=False)
    ctx = qlib.Context(device=qlib.DEVICE_NONE)
    ctx.set_expansion(ex)
    pr, pa = qlib.topologies(1, 1)
    ctx.set_topologies(0, qlib.MODE_BOLD, 1, 1, pr, pa)
    with pytest.raises(qlib.QiwError) as ei:
        ctx.bsize = 2
        ctx.eval(0.0, 1.0, 2.0, [0], 16)
    assert ei.value.code == 2
    with pytest.raises(qlib.QiwError):
        ctx.sobol_points(qlib.sobol_direction_numbers(2), None, 0, 4)


def test_sobol_host_bit_exact(qlib, oracle_lib):
    for D in (0, 1, 2, 5, 12, 16, 40):
        assert np.array_equal(qlib.sobol_direction_numbers(D), oracle_lib.sobol_direction_numbers(D))
    with pytest.raises(ValueError):
        qlib.sobol_direction_numbers(65)
    rng = np.random.default_rng(5)
    for D in (1, 4, 12):
        m = qlib.sobol_direction_numbers(D)
        sb, lb = rng.integers(0, 2, (D, 32)), rng.integers(0, 2, (D, 32, 32))
        a, b = qlib.sobol_scramble(m, sb, lb), oracle_lib.sobol_scramble(m, sb, lb)
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_topologies_bit_exact(qlib, oracle_lib):
    for n in range(0, 6):
        for k in [None] + list(range(0, 2 * n)):
            for ext in (False, True):
                a, b = qlib.topologies(n, k, ext), oracle_lib.topologies(n, k, ext)
                assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]), (n, k, ext)
    # test/diagrammatics.jl:53-60 irreducible counts, (2n-1)!! totals (:26-42); order 7 is SURVEY N1's target size
    assert [len(qlib.topologies(n, 1)[1]) for n in range(1, 8)] == [1, 1, 4, 27, 248, 2830, 38232]
    assert [len(qlib.topologies(n)[1]) for n in range(8)] == [1, 1, 3, 15, 105, 945, 10395, 135135]
    a, b = qlib.topologies(6, 5), oracle_lib.topologies(6, 5)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_partitioning(qlib, oracle_lib):
    from qinchworm_b200 import mpi
    assert mpi.split_count(10, 3) == [4, 3, 3]
    assert list(mpi.range_from_chunks_and_idx([4, 3, 3], 2)) == [5, 6, 7]
    for N in (0, 1, 7, 10, 1024, 2 ** 20 + 3):
        for R in (1, 2, 3, 8):
            tot = 0
            for r in range(R):
                s, c = qlib.rank_sub_range(N, R, r)
                assert (s, c) == oracle_lib.rank_sub_range(N, R, r)
                rr = mpi.rank_sub_range(N, r, R)
                assert (rr.start - 1, len(rr)) == (s, c)
                assert s == tot
                tot += c
            assert tot == N


@pytest.mark.parametrize("shared", ["0", "1"])
@pytest.mark.parametrize("model", ["anderson", "single_level", "dimer"])
def test_compiled_programs_replay_to_oracle(qlib, oracle_lib, model, shared, monkeypatch):
    """qiw_set_topologies (host compiler) vs the oracle's recursive evaluator, all three modes; the lane program both
    without and with records shared by two initial sectors (QIW_LANE_DUAL; the default turns them on at order 5)."""
    from qinchworm_b200.expansion import add_corr_operators
    monkeypatch.setenv("QIW_LANE_DUAL", shared)
    n_shared = 0
    rng = np.random.default_rng(11)
    if model == "anderson":
        ex, grid, f = models.anderson(n_tau=30, corr=True)
        max_order = 3
    elif model == "single_level":
        ex, grid, f = models.single_level(n_tau=16, spline=True)
        add_corr_operators(ex, (f.c("0"), f.c_dag("0")))
        max_order = 4
    else:
        ex, grid, f = models.hubbard_dimer_impurity(n_tau=16)
        add_corr_operators(ex, (f.c(1), f.c_dag(1)))
        max_order = 3
    ex.P = ex.P * (1 + 0.1 * rng.random(ex.P.shape))
    pl = ex.flatten()
    ctx = qlib.Context(device=qlib.DEVICE_NONE)
    ctx.set_expansion(ex)
    o = oracle_lib.Oracle(pl, ex.P)
    tau = grid.tau
    eid = 0
    for mode in (qlib.MODE_BOLD, qlib.MODE_BARE, qlib.MODE_CORR):
        for corr in range(len(ex.corr_operators) if mode == qlib.MODE_CORR else 1):
            for order in range(0, max_order + 1):
                ks = [None] if mode == qlib.MODE_BARE else ([0] if order == 0 else range(1, 2 * order))
                for k in ks:
                    pr, pa = qlib.topologies(order, None if mode == qlib.MODE_BARE else k, mode == qlib.MODE_CORR)
                    if len(pa) == 0 or (order == max_order and k not in (None, 1, 2 * order - 1)):
                        continue
                    kk = 2 * order if mode == qlib.MODE_BARE else k
                    ctx.set_topologies(eid, mode, order, kk, pr, pa, corr_idx=corr)
                    o.set_topologies(eid, mode, order, kk, pr, pa)
                    st, prog = ctx.entry_stats(eid), ctx.entry_program(eid)
                    if mode == qlib.MODE_BARE:
                        t_i, t_w, t_f = 0.0, tau[0], tau[1]
                    elif mode == qlib.MODE_BOLD:
                        t_i, t_w, t_f = 0.0, tau[7], tau[8]
                    else:
                        t_i, t_w, t_f = 0.0, tau[9], tau[-1]
                    times = np.zeros((2, 2 * order))
                    for i in range(2):
                        if mode == qlib.MODE_BARE:
                            times[i] = np.sort(rng.uniform(t_i, t_f, 2 * order))[::-1]
                        else:
                            times[i, :kk] = np.sort(rng.uniform(t_w, t_f, kk))[::-1]
                            times[i, kk:] = np.sort(rng.uniform(t_i, t_w, 2 * order - kk))[::-1]
                    ref = o.eval_at_times(eid, t_i, t_w, t_f, times, corr_idx=corr)
                    fl, lv = o.last_counts()
                    got = np.array([run_program(prog, ex, pl, mode, t_i, t_w, t_f, times[i]) for i in range(2)])
                    assert np.abs(got - ref).max() <= 1e-9 * max(np.abs(ref).max(), 1e-300)  # interpreter uses SciPy splines
                    rec = ctx.entry_records(eid)     # the factorised records the kernel actually executes
                    assert rec["n_leaves"] == lv and rec["L2"] == rec["K"] + order
                    got2 = np.array([run_records(rec, prog, ex, pl, mode, t_i, t_w, t_f, times[i]) for i in range(2)])
                    assert np.abs(got2 - ref).max() <= 1e-9 * max(np.abs(ref).max(), 1e-300)
                    lp = ctx.entry_lane_program(eid)   # ... and the lane program the step kernel executes (lane = sample)
                    n_shared += sum(int(n) for sc, _, n, _ in lp["sections"] if int(sc) >> 16)
                    assert shared == "1" or all(int(sc) >> 16 == 0 for sc, _, _, _ in lp["sections"])
                    res3 = [run_lane_program(lp, prog, ex, pl, mode, t_i, t_w, t_f, times[i]) for i in range(2)]
                    assert all(nm == lv for _, nm in res3)
                    got3 = np.array([r for r, _ in res3])
                    assert np.abs(got3 - ref).max() <= 1e-9 * max(np.abs(ref).max(), 1e-300)
                    assert st["n_leaves"] == lv and st["flops_per_sample"] == fl and st["n_top"] == len(pa)
                    eid += 1
    assert eid > 10
    assert shared == "0" or model == "single_level" or n_shared > 0      # (two sectors that never share a Delta set there)


@pytest.mark.parametrize("unit_cost", [None, "150"])
@pytest.mark.parametrize("model", ["two_level_mixed", "three_orbital", "two_band"])
def test_walk_units_replay_to_oracle(qlib, oracle_lib, model, unit_cost, monkeypatch):
    """Sector blocks larger than 1x1: the walk units block_walk_kernel executes (sub-trees of bounded cost, one
    unit per group of at most two columns; built on the host by qiw_set_topologies) replayed in numpy equal the
    oracle's recursive evaluator, all three modes."""
    from program_interp import run_walk_units
    if unit_cost:      # cost bound of a unit (default max(4000, entry cost / 1024)): a small one cuts every tree deeply
        monkeypatch.setenv("QIW_WALK_UNIT_COST", unit_cost)
    rng = np.random.default_rng(5)
    if model == "two_level_mixed":
        ex, grid, f = models.two_level_mixed(n_tau=12, theta=0.6)
        max_order = 3
    elif model == "three_orbital":
        ex, grid, f = models.three_orbital(n_tau=12)
        max_order = 2
    else:
        ex, grid, f = models.two_band(n_tau=10)
        max_order = 2
    ex.P = ex.P * (1 + 0.1 * rng.random(ex.P.shape))
    pl = ex.flatten()
    ctx = qlib.Context(device=qlib.DEVICE_NONE)
    ctx.set_expansion(ex)
    o = oracle_lib.Oracle(pl, ex.P)
    tau = grid.tau
    eid, n_units, n_split = 0, 0, 0
    for mode in (qlib.MODE_BOLD, qlib.MODE_BARE, qlib.MODE_CORR):
        for order in range(0, max_order + 1):
            ks = [None] if mode == qlib.MODE_BARE else ([0] if order == 0 else (1, 2 * order - 1))
            for k in ks:
                pr, pa = qlib.topologies(order, None if mode == qlib.MODE_BARE else k, mode == qlib.MODE_CORR)
                if len(pa) == 0:
                    continue
                kk = 2 * order if mode == qlib.MODE_BARE else k
                ctx.set_topologies(eid, mode, order, kk, pr, pa)
                o.set_topologies(eid, mode, order, kk, pr, pa)
                prog, units = ctx.entry_program(eid), ctx.entry_walk_units(eid)
                if mode == qlib.MODE_BARE:
                    t_i, t_w, t_f = 0.0, tau[0], tau[1]
                elif mode == qlib.MODE_BOLD:
                    t_i, t_w, t_f = 0.0, tau[5], tau[6]
                else:
                    t_i, t_w, t_f = 0.0, tau[4], tau[-1]
                times = np.zeros(2 * order)
                if mode == qlib.MODE_BARE:
                    times[:] = np.sort(rng.uniform(t_i, t_f, 2 * order))[::-1]
                else:
                    times[:kk] = np.sort(rng.uniform(t_w, t_f, kk))[::-1]
                    times[kk:] = np.sort(rng.uniform(t_i, t_w, 2 * order - kk))[::-1]
                ref = o.eval_at_times(eid, t_i, t_w, t_f, times[None, :])[0]
                got = run_walk_units(units, prog, ex, pl, mode, t_i, t_w, t_f, times)
                assert np.abs(got - ref).max() <= 1e-9 * max(np.abs(ref).max(), 1e-300), (mode, order, k)
                n_units += len(units["unit_off"]) - 1
                n_split += (len(units["unit_off"]) - 1) > (len(prog["tree_off"]) - 1)
                eid += 1
    assert eid >= 8 and n_split > 0     # some entries were cut into more units than trees
    if unit_cost:
        assert n_units > 20 * eid


def test_offdiagonal_block_is_an_error(qlib):
    """Under-resolved sectors must be reported like the reference's @assert
    (src/topology_eval.jl:462), at compile time."""
    from qinchworm_b200.ed import EDCore, FockSpace
    from qinchworm_b200.expansion import Expansion, InteractionPair
    from qinchworm_b200.gf import ImaginaryTimeGrid, delta_dos_gf, ph_conj
    f = FockSpace([["a"], ["b"]])
    H = 0.3 * f.n_op("a") + 0.5 * f.n_op("b")
    grid = ImaginaryTimeGrid(2.0, 8)
    D = delta_dos_gf(grid, 0.4)
    ex = Expansion(EDCore(f, H), grid, [InteractionPair(f.c_dag("a"), f.c("b"), D),
                                         InteractionPair(f.c("b"), f.c_dag("a"), ph_conj(D))])
    ctx = qlib.Context(device=qlib.DEVICE_NONE)
    ctx.set_expansion(ex)
    pr, pa = qlib.topologies(1, 1)
    with pytest.raises(qlib.QiwError) as ei:
        ctx.set_topologies(0, qlib.MODE_BOLD, 1, 1, pr, pa)
    assert ei.value.code == 4


def test_readme_work_counts(qlib):
    """Work counters of the README configuration: 280 bold topologies, 3 492 surviving
    configurations over the 17 bold entries, 2 656 for bare order 4."""
    ex, grid, f = models.anderson(n_tau=20)
    ctx = qlib.Context(device=qlib.DEVICE_NONE)
    ctx.set_expansion(ex)
    leaves = tops = 0
    eid = 0
    for order in range(0, 5):
        for k in ([0] if order == 0 else range(1, 2 * order)):
            pr, pa = qlib.topologies(order, k)
            ctx.set_topologies(eid, qlib.MODE_BOLD, order, k, pr, pa)
            st = ctx.entry_stats(eid)
            leaves += st["n_leaves"]
            tops += st["n_top"] if order else 0
            eid += 1
    assert eid == 17 and tops == 280 and leaves == 3492
    pr, pa = qlib.topologies(4)
    ctx.set_topologies(eid, qlib.MODE_BARE, 4, 8, pr, pa)
    assert ctx.entry_stats(eid)["n_leaves"] == 2656


def _digest(obj, h):
    if isinstance(obj, dict):
        for k in sorted(obj):
            h.update(str(k).encode())
            _digest(obj[k], h)
    elif isinstance(obj, np.ndarray):
        h.update(str(obj.shape).encode())
        h.update(np.ascontiguousarray(obj).tobytes())
    elif isinstance(obj, (list, tuple)):
        for x in obj:
            _digest(x, h)
    else:
        h.update(repr(obj).encode())


@pytest.mark.parametrize("model", ["anderson", "two_band"])
def test_threaded_compile_is_bit_identical(qlib, model, monkeypatch):
    """qiw_set_topologies walks big entries one initial-sector group per host thread and renumbers the fragments'
    coefficients and pair-interaction slots into the global first-seen order: the program, the factorised records / lane program and the walk units must equal the sequential compilation (QIW_COMPILE_THREADS=1) bit for bit."""
    import hashlib
    if model == "anderson":
        ex, grid, f = models.anderson(n_tau=20, corr=True)
        cases = [(qlib.MODE_BOLD, 5, 1, 0), (qlib.MODE_BOLD, 5, 6, 0), (qlib.MODE_BARE, 5, 10, 0), (qlib.MODE_CORR, 5, 3, 1),
                 (qlib.MODE_BOLD, 6, 11, 0)]
    else:
        ex, grid, f = models.two_band(n_tau=8)
        cases = [(qlib.MODE_BOLD, 3, 1, 0), (qlib.MODE_BOLD, 3, 4, 0), (qlib.MODE_CORR, 3, 2, 0), (qlib.MODE_BARE, 3, 6, 0)]
        monkeypatch.setenv("QIW_WALK_UNIT_COST", "150")     # deep cuts: many units per tree, long replayed paths
    digests = {}
    for threads in ("1", "8"):
        monkeypatch.setenv("QIW_COMPILE_THREADS", threads)
        ctx = qlib.Context(device=qlib.DEVICE_NONE)
        ctx.set_expansion(ex)
        h = hashlib.sha256()
        for eid, (mode, order, k, corr) in enumerate(cases):
            pr, pa = qlib.topologies(order, None if mode == qlib.MODE_BARE else k, mode == qlib.MODE_CORR)
            assert len(pa) > 0
            ctx.set_topologies(eid, mode, order, k, pr, pa, corr_idx=corr)
            _digest(ctx.entry_stats(eid), h)
            _digest(ctx.entry_program(eid), h)
            if model == "anderson":
                _digest(ctx.entry_records(eid), h)
                _digest(ctx.entry_lane_program(eid), h)
            else:
                _digest(ctx.entry_walk_units(eid), h)
        digests[threads] = h.hexdigest()
        ctx.close()
    assert digests["1"] == digests["8"]


def test_host_stepped_step_bookkeeping(qlib):
    """The host-stepped loop's fast path (default RandomizationParams: long-lived id/result buffers, per-order sums as
    one matmul) returns what the general path (eval_entries + _order_sums) returns, and normalize_at is the
    reference's normalize! (src/ppgf.jl:646-668): lambda from the largest diagonal element, every row times
    exp(-lambda tau_k).  The library call is replaced by a stub that fills the result buffer (no GPU here)."""
    from qinchworm_b200 import lib, ppgf
    from qinchworm_b200.inchworm import RandomizationParams, Solver, _bold_entries, _order_sums, inchworm_step
    ex, grid = models.anderson(n_tau=16)[:2]
    ctx = lib.Context(device=lib.DEVICE_NONE)
    solver = Solver(ex, ctx=ctx)
    top = _bold_entries(solver, range(0, 4), 64, None, None)
    rng = np.random.default_rng(7)
    res = rng.standard_normal((len(top), ctx.bsize)) + 1j * rng.standard_normal((len(top), ctx.bsize))
    seen = {}

    def eval_prepared(t_i, t_w, t_f, n, ids_ptr, N, out_ptr):
        seen["args"] = (t_i, t_w, t_f, n, N)
        seen["ids"] = np.ctypeslib.as_array(ids_ptr, shape=(n,)).copy()
        np.ctypeslib.as_array(out_ptr, shape=(n * ctx.bsize * 2,))[:] = res.view(np.float64).reshape(-1)

    ctx.eval_prepared = eval_prepared
    total, contribs, contribs_std = inchworm_step(solver, grid, 0, 3, 4, top)
    assert seen["args"] == (grid.tau[0], grid.tau[3], grid.tau[4], len(top), 64)
    assert list(seen["ids"]) == [td.entry_id for td in top]
    std = np.full_like(res, np.nan)
    std[[j for j, td in enumerate(top) if td.order == 0]] = 0.0
    t2, c2, s2 = _order_sums(top, res, std, ctx.bsize)
    assert np.allclose(total, t2, rtol=1e-14, atol=1e-14)
    assert sorted(contribs) == sorted(c2) == [0, 1, 2, 3]
    for o in c2:
        assert np.allclose(contribs[o], c2[o], rtol=1e-14, atol=1e-14)
        assert np.array_equal(np.isnan(contribs_std[o]), np.isnan(s2[o])) and (o > 0 or np.all(contribs_std[o] == 0))
    # a second call reuses the prepared buffers and must not alias the first call's results
    res2 = res * 2.0
    res_saved, res = res, res2
    total_b, contribs_b, _ = inchworm_step(solver, grid, 0, 4, 5, top)
    assert np.allclose(total_b, 2.0 * t2) and np.allclose(total, t2)
    assert len(solver._steps) == 1
    # the same compiled entries with another sample count (a Solver is reused across calls): the count of THIS call travels
    top2 = _bold_entries(solver, range(0, 4), 256, None, None)
    assert [td.entry_id for td in top2] == [td.entry_id for td in top]
    inchworm_step(solver, grid, 0, 4, 5, top2)
    assert seen["args"][4] == 256 and len(solver._steps) == 1
    # a randomised call takes the general path (it needs the GPU): only check that the fast path is not chosen
    assert RandomizationParams().rng is None and RandomizationParams().N_seqs == 1
    # normalize_at
    P0 = -1j * (0.5 + rng.random(ex.P.shape))
    ex.P[:] = P0
    lam = ppgf.normalize_at(ex, 5)
    diag = ppgf._diag_indices(ex)
    lam_ref = np.log(np.max(-P0[5, diag].imag)) / grid.tau[5]
    assert abs(lam - lam_ref) < 1e-14 * max(1.0, abs(lam_ref))
    assert np.allclose(ex.P, P0 * np.exp(-grid.tau * lam_ref)[:, None], rtol=1e-14, atol=0)
    assert abs(np.max(-ex.P[5, diag].imag) - 1.0) < 1e-13


def test_order5_lane_program_shares_records_by_default(qlib, oracle_lib, monkeypatch):
    """From order 5 on the lane program uses records shared by two initial sectors without being asked to
    (QIW_LANE_DUAL unset); replayed on CPU against the oracle for one bold order-5 entry."""
    monkeypatch.delenv("QIW_LANE_DUAL", raising=False)
    rng = np.random.default_rng(5)
    ex, grid, f = models.anderson(n_tau=30)
    ex.P = ex.P * (1 + 0.1 * rng.random(ex.P.shape))
    pl = ex.flatten()
    ctx = qlib.Context(device=qlib.DEVICE_NONE)
    ctx.set_expansion(ex)
    o = oracle_lib.Oracle(pl, ex.P)
    tau = grid.tau
    for eid, (order, k) in enumerate([(4, 3), (5, 3)]):
        pr, pa = qlib.topologies(order, k)
        ctx.set_topologies(eid, qlib.MODE_BOLD, order, k, pr, pa)
        o.set_topologies(eid, qlib.MODE_BOLD, order, k, pr, pa)
        lp, prog = ctx.entry_lane_program(eid), ctx.entry_program(eid)
        n_shared = sum(int(n) for sc, _, n, _ in lp["sections"] if int(sc) >> 16)
        assert (n_shared > 0) == (order >= 5)
        t_i, t_w, t_f = 0.0, tau[11], tau[12]
        times = np.zeros((1, 2 * order))
        times[0, :k] = np.sort(rng.uniform(t_w, t_f, k))[::-1]
        times[0, k:] = np.sort(rng.uniform(t_i, t_w, 2 * order - k))[::-1]
        ref = o.eval_at_times(eid, t_i, t_w, t_f, times)
        got, n_members = run_lane_program(lp, prog, ex, pl, qlib.MODE_BOLD, t_i, t_w, t_f, times[0])
        assert n_members == o.last_counts()[1]
        assert np.abs(got - ref[0]).max() <= 1e-9 * np.abs(ref).max()
