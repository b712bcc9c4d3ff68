"""GPU parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same
seeded inputs, and against the reference's golden vectors.  Tolerance: FP64 relative 1e-10
(BASELINE.json north_star); Sobol points bit-exact."""
import numpy as np
import pytest

import models
from conftest import load_golden

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def relerr(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(b).max(), 1e-300)


def relerr_elem(a, b, floor=1e-3):
    """Element-wise relative error: every component is compared with its OWN magnitude (components below
    `floor` times the largest one are compared with that floor), so that small sectors — e.g. the doubly
    occupied one of the Anderson model, 2 % of the largest — are held to the same 1e-10 as the large ones."""
    a, b = np.asarray(a), np.asarray(b)
    scale = np.maximum(np.abs(b), floor * max(np.abs(b).max(), 1e-300))
    return float((np.abs(a - b) / scale).max())


def test_sobol_points_bit_exact(gpu_ctx, qlib, oracle_lib):
    for D in (1, 5, 12):
        m = qlib.sobol_direction_numbers(D)
        x0 = np.zeros(D, dtype=np.uint32)
        ref, _ = oracle_lib.sobol_points(m, x0, 0, 4096)
        got = gpu_ctx.sobol_points(m, x0, 0, 4096)
        assert np.array_equal(got, ref)
        ref, _ = oracle_lib.sobol_points(m, x0, 1000, 77)       # skip!(seq, 1000, exact=true)
        assert np.array_equal(gpu_ctx.sobol_points(m, x0, 1000, 77), ref)
    # scrambled sequence with the reference's MockRNG bits (test/scrambled_sobol.jl:97-105)
    G = load_golden("sobol_tables.json")

    def mock_bits(shape):
        v = np.array([bin(i).count("1") % 2 for i in range(int(np.prod(shape)))], dtype=np.uint8)
        return v.reshape(shape, order="F")
    for D, key in ((1, "scrambled_D1"), (5, "scrambled_D5")):
        m, x0 = qlib.sobol_scramble(qlib.sobol_direction_numbers(D), mock_bits((D, 32)), mock_bits((D, 32, 32)))
        pts = gpu_ctx.sobol_points(m, x0, 0, 8).astype(np.float64) * 2.0 ** -32
        assert np.abs(pts - np.array(G[key])).max() < 1e-10


def test_topology_eval_golden(gpu_ctx, qlib):
    """test/topology_eval.jl:137-151 through qiw_eval_at_times."""
    G = load_golden("topology_eval_h5.json")
    ex, grid, f = models.single_level(n_tau=30, spline=False, rev="transpose")
    gpu_ctx.set_expansion(ex)
    pairs, parity = qlib.topologies(3, 1)
    assert len(parity) == 4
    gpu_ctx.set_topologies(0, qlib.MODE_BOLD, 3, 1, pairs, parity)
    tau = grid.tau
    tw, tf = tau[6], tau[7]
    times = np.zeros((100, 6))
    times[:, 0] = tw + (tf - tw) * G["/x1_list"]
    times[:, 1:] = tw * G["/xs_list"]
    got = gpu_ctx.eval_at_times(0, 0.0, tw, tf, times)
    ref = G["/values"].T[:, ::-1]   # golden sector 1 = occupied = our sector 1
    assert relerr(got, ref) < RTOL


@pytest.mark.parametrize("name", ["single_level", "anderson"])
def test_step_vs_oracle(gpu_ctx, qlib, oracle_lib, name):
    """One bold step, one bare step and one correlator point against the oracle."""
    if name == "single_level":
        ex, grid, f = models.single_level(n_tau=20, spline=True)
        from qinchworm_b200.expansion import add_corr_operators
        add_corr_operators(ex, (f.c("0"), f.c_dag("0")))
        orders, N = range(0, 4), 2 ** 8
    else:
        ex, grid, f = models.anderson(n_tau=50, corr=True)
        orders, N = range(0, 4), 2 ** 9
    rng = np.random.default_rng(7)
    ex.P = ex.P * (1.0 + 0.05 * rng.random(ex.P.shape))   # away from the atomic limit
    pl = gpu_ctx.set_expansion(ex)
    o = oracle_lib.Oracle(pl, ex.P)
    tau = grid.tau
    eid = 0
    for mode, (ki, kw, kf), corr in ((qlib.MODE_BOLD, (0, 11, 12), 0), (qlib.MODE_BARE, (0, 0, 1), 0),
                                     (qlib.MODE_CORR, (0, 9, len(tau) - 1), 0),
                                     (qlib.MODE_CORR, (0, 5, len(tau) - 1), len(ex.corr_operators) - 1)):
        ids = []
        for order in orders:
            ks = [None] if mode == qlib.MODE_BARE else ([0] if order == 0 else range(1, 2 * order))
            for k in ks:
                pr, pa = qlib.topologies(order, None if mode == qlib.MODE_BARE else k, mode == qlib.MODE_CORR)
                if len(pa) == 0:
                    continue
                kk = 2 * order if mode == qlib.MODE_BARE else k
                gpu_ctx.set_topologies(eid, mode, order, kk, pr, pa, corr_idx=corr)
                o.set_topologies(eid, mode, order, kk, pr, pa)
                ids.append(eid)
                eid += 1
        got = gpu_ctx.eval(tau[ki], tau[kw], tau[kf], ids, N)
        ref = o.eval(tau[ki], tau[kw], tau[kf], ids, N, corr_idx=corr)
        assert relerr(got, ref) < RTOL, (mode, corr, relerr(got, ref))
        # ragged sample range + scrambled sequence, no all-reduce
        sob = []
        for e in ids:
            D = 2 * gpu_ctx.entry_order[e]
            m = qlib.sobol_direction_numbers(D)
            if D:
                m, x0 = qlib.sobol_scramble(m, rng.integers(0, 2, (D, 32)), rng.integers(0, 2, (D, 32, 32)))
            else:
                x0 = np.zeros(0, dtype=np.uint32)
            sob.append((m, x0))
        got = gpu_ctx.eval_range(tau[ki], tau[kw], tau[kf], ids, N, 37, 101, sobol=sob)
        ref = o.eval(tau[ki], tau[kw], tau[kf], ids, N, start=37, count=101, corr_idx=corr, sobol=sob)
        assert relerr(got, ref) < RTOL, ("range", mode, corr, relerr(got, ref))


def test_inchworm_golden(gpu_ctx, qlib):
    """test/inchworm.jl:179-212: full inchworm! + correlator_2p against test/inchworm.h5."""
    from qinchworm_b200.expansion import add_corr_operators
    from qinchworm_b200.inchworm import Solver, correlator_2p, inchworm
    G = load_golden("inchworm_h5.json")
    ex, grid, f = models.single_level(n_tau=20, spline=True)
    solver = Solver(ex, ctx=gpu_ctx)
    inchworm(ex, grid, range(0, 4), range(0, 3), 2 ** 8, solver=solver, device_resident=False)   # host-stepped loop
    assert relerr(ex.P[:, 1], G["/inchworm/P/1"].ravel()) < RTOL
    assert relerr(ex.P[:, 0], G["/inchworm/P/2"].ravel()) < RTOL
    add_corr_operators(ex, (f.c("0"), f.c_dag("0")))
    g = -correlator_2p(ex, grid, range(0, 4), 2 ** 8, solver=solver)[0]
    assert relerr(g, G["/inchworm/g"].ravel()) < RTOL


def test_hubbard_dimer_physics(gpu_ctx, qlib):
    """test/dimers.jl:124-196: density matrix of the Hubbard dimer vs exact ED (< 1e-4)."""
    from qinchworm_b200 import ppgf
    from qinchworm_b200.inchworm import Solver, inchworm
    ex, grid, f = models.hubbard_dimer_impurity(n_tau=32)
    inchworm(ex, grid, range(0, 3), range(0, 3), 8 * 2 ** 5, solver=Solver(ex, ctx=gpu_ctx))
    ppgf.normalize(ex)
    rho = ex.ed.to_fock_basis(ppgf.density_matrix(ex))
    ref = models.hubbard_dimer_exact_rho()
    assert np.abs(rho - ref).max() < 1e-4


def test_inchworm_device_resident(gpu_ctx, qlib):
    """qiw_inchworm_run (set_ppgf!/normalize! on the device) against test/inchworm.h5 and against
    the host-driven loop, including the order-resolved contributions."""
    from qinchworm_b200.inchworm import Solver, inchworm
    G = load_golden("inchworm_h5.json")
    ex, grid, f = models.single_level(n_tau=20, spline=True)
    Po, _ = inchworm(ex, grid, range(0, 4), range(0, 3), 2 ** 8, solver=Solver(ex, ctx=gpu_ctx), device_resident=True)
    assert relerr(ex.P[:, 1], G["/inchworm/P/1"].ravel()) < RTOL
    assert relerr(ex.P[:, 0], G["/inchworm/P/2"].ravel()) < RTOL
    ex2, grid2, _ = models.single_level(n_tau=20, spline=True)
    Po2, _ = inchworm(ex2, grid2, range(0, 4), range(0, 3), 2 ** 8, solver=Solver(ex2, ctx=gpu_ctx), device_resident=False)
    assert relerr(ex.P, ex2.P) < RTOL
    for o in Po:
        assert relerr(Po[o], Po2[o]) < RTOL


def test_readme_anderson_run(gpu_ctx, qlib, oracle_lib):
    """README.md:34-158 configuration at reduced n_tau (the oracle finishes in seconds): P(tau), Z and
    rho_imp of the device-resident run vs the oracle, tolerance 1e-10; PH-symmetric variant gives
    rho = 1/4 (test/bethe.jl:104-126)."""
    from qinchworm_b200 import ppgf
    from qinchworm_b200.inchworm import Solver, inchworm
    ex, grid, f = models.anderson(n_tau=24)
    ref = oracle_lib.inchworm(ex.flatten(), ex.P, range(0, 5), range(0, 5), 2 ** 7)["P"]
    inchworm(ex, grid, range(0, 5), range(0, 5), 2 ** 7, solver=Solver(ex, ctx=gpu_ctx), device_resident=True)
    assert relerr(ex.P, ref) < RTOL
    Z = ppgf.partition_function(ex)
    Zref = (1j * ref[-1]).sum()
    assert abs(Z - Zref) < RTOL * abs(Zref)
    # particle-hole symmetric point: eps = -U/2
    ex, grid, f = models.anderson(n_tau=24, eps=-0.5, U=1.0)
    inchworm(ex, grid, range(0, 3), range(0, 3), 2 ** 8, solver=Solver(ex, ctx=gpu_ctx), device_resident=True)
    ppgf.normalize(ex)
    rho = np.array([d[0, 0].real for d in ppgf.density_matrix(ex)])
    assert np.abs(rho[1] - rho[2]) < 1e-12 and abs(rho.sum() - 1) < 1e-12
    assert np.abs(rho[0] - rho[3]) < 2e-3


@pytest.mark.parametrize("name", ["two_level_mixed", "three_orbital", "two_band"])
def test_block_models_vs_oracle(gpu_ctx, qlib, oracle_lib, name):
    """Sector blocks larger than 1x1 (unpinned at the reference level; the oracle is
    self-validated by basis-rotation invariance): bold / bare / correlator entries against the oracle,
    per-sample values through qiw_eval_at_times, and a full device-resident run."""
    from qinchworm_b200.inchworm import Solver, inchworm
    if name == "two_level_mixed":
        ex, grid, f = models.two_level_mixed(n_tau=12, theta=0.6)
        orders, N = range(0, 4), 2 ** 7
    elif name == "three_orbital":
        ex, grid, f = models.three_orbital(n_tau=12)
        assert sorted(ex.dims) == [1, 1, 3, 3]
        orders, N = range(0, 4), 2 ** 7
    else:
        ex, grid, f = models.two_band(n_tau=10)
        assert sorted(ex.dims) == [1, 1, 1, 1, 2, 2, 2, 2, 4]
        orders, N = range(0, 3), 2 ** 6
    rng = np.random.default_rng(3)
    ex.P = ex.P * (1.0 + 0.05 * rng.random(ex.P.shape))
    pl = gpu_ctx.set_expansion(ex)
    o = oracle_lib.Oracle(pl, ex.P)
    tau = grid.tau
    eid = 0
    for mode, (ki, kw, kf) in ((qlib.MODE_BOLD, (0, 5, 6)), (qlib.MODE_BARE, (0, 0, 1)), (qlib.MODE_CORR, (0, 4, len(tau) - 1))):
        ids = []
        for order in orders:
            ks = [None] if mode == qlib.MODE_BARE else ([0] if order == 0 else range(1, 2 * order))
            for k in ks:
                pr, pa = qlib.topologies(order, None if mode == qlib.MODE_BARE else k, mode == qlib.MODE_CORR)
                if len(pa) == 0:
                    continue
                kk = 2 * order if mode == qlib.MODE_BARE else k
                gpu_ctx.set_topologies(eid, mode, order, kk, pr, pa)
                o.set_topologies(eid, mode, order, kk, pr, pa)
                st = gpu_ctx.entry_stats(eid)
                ids.append(eid)
                eid += 1
        got = gpu_ctx.eval(tau[ki], tau[kw], tau[kf], ids, N)
        ref = o.eval(tau[ki], tau[kw], tau[kf], ids, N)
        assert relerr(got, ref) < RTOL, (mode, relerr(got, ref))
        got = gpu_ctx.eval_range(tau[ki], tau[kw], tau[kf], ids, N, 5, 33)
        ref = o.eval(tau[ki], tau[kw], tau[kf], ids, N, start=5, count=33)
        assert relerr(got, ref) < RTOL, ("range", mode, relerr(got, ref))
    # per-sample evaluator values for one order-2 bold entry
    pr, pa = qlib.topologies(2, 2)
    gpu_ctx.set_topologies(eid, qlib.MODE_BOLD, 2, 2, pr, pa)
    o.set_topologies(eid, qlib.MODE_BOLD, 2, 2, pr, pa)
    times = np.zeros((7, 4))
    for i in range(7):
        times[i, :2] = np.sort(rng.uniform(tau[5], tau[6], 2))[::-1]
        times[i, 2:] = np.sort(rng.uniform(0, tau[5], 2))[::-1]
    got = gpu_ctx.eval_at_times(eid, 0.0, tau[5], tau[6], times)
    ref = o.eval_at_times(eid, 0.0, tau[5], tau[6], times)
    assert relerr(got, ref) < RTOL
    fl, lv = o.last_counts()
    st = gpu_ctx.entry_stats(eid)
    assert st["n_leaves"] == lv and st["flops_per_sample"] == fl
    # full run on the device vs the oracle's driver
    ex2 = {"two_level_mixed": lambda: models.two_level_mixed(n_tau=12, theta=0.6), "three_orbital": lambda: models.three_orbital(n_tau=12),
           "two_band": lambda: models.two_band(n_tau=10)}[name]()[0]
    refP = oracle_lib.inchworm(ex2.flatten(), ex2.P, range(0, 3), range(0, 3), 2 ** 5)["P"]
    inchworm(ex2, ex2.grid, range(0, 3), range(0, 3), 2 ** 5, solver=Solver(ex2, ctx=gpu_ctx), device_resident=True)
    assert relerr(ex2.P, refP) < RTOL


def test_real_and_complex_arithmetic_agree(gpu_ctx, qlib, oracle_lib, monkeypatch):
    """The real-arithmetic fast path (all operands purely imaginary-time) and the general complex path
    of the step kernel give the same numbers; a complex hybridisation forces the complex path and still
    matches the oracle."""
    ex, grid, f = models.anderson(n_tau=40)
    rng = np.random.default_rng(11)
    ex.P = ex.P * (1.0 + 0.05 * rng.random(ex.P.shape))
    pl = gpu_ctx.set_expansion(ex)
    o = oracle_lib.Oracle(pl, ex.P)
    ids = []
    for order in range(0, 5):
        for k in ([0] if order == 0 else range(1, 2 * order)):
            pr, pa = qlib.topologies(order, k)
            gpu_ctx.set_topologies(len(ids), qlib.MODE_BOLD, order, k, pr, pa)
            o.set_topologies(len(ids), qlib.MODE_BOLD, order, k, pr, pa)
            ids.append(len(ids))
    tau = grid.tau
    N = 2 ** 9
    ref = o.eval(0.0, tau[20], tau[21], ids, N)
    got_real = gpu_ctx.eval(0.0, tau[20], tau[21], ids, N)
    monkeypatch.setenv("QIW_FORCE_COMPLEX", "1")
    got_cplx = gpu_ctx.eval(0.0, tau[20], tau[21], ids, N)
    monkeypatch.delenv("QIW_FORCE_COMPLEX")
    assert relerr(got_real, ref) < RTOL and relerr(got_cplx, ref) < RTOL
    assert relerr(got_real, got_cplx) < 1e-13
    # a P table with a real part switches the library to complex arithmetic by itself
    P2 = ex.P * (1.0 + 0.02j)
    gpu_ctx.set_P(0, P2)
    o2 = oracle_lib.Oracle(pl, P2)
    eid = 0
    for order in range(0, 5):
        for k in ([0] if order == 0 else range(1, 2 * order)):
            pr, pa = qlib.topologies(order, k)
            o2.set_topologies(eid, qlib.MODE_BOLD, order, k, pr, pa)
            eid += 1
    got = gpu_ctx.eval(0.0, tau[20], tau[21], ids, N)
    ref2 = o2.eval(0.0, tau[20], tau[21], ids, N)
    assert relerr(got, ref2) < RTOL


def test_multi_gpu_allreduce(qlib):
    """One rank per GPU (torchrun, 2 ranks): sharded Sobol ranges + the all-reduce (peer-memory exchange
    inside the step kernel, and the NCCL fallback) against the single-process oracle.  Needs 2 GPUs."""
    import subprocess
    import sys
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(here, "multigpu_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "multigpu_check OK" in out.stdout, out.stdout[-2000:] + out.stderr[-3000:]


def test_correlator_batched_vs_stepped(gpu_ctx, qlib, oracle_lib):
    """bench/bethe_gf_convergence (C2): correlator_2p with all grid points in one qiw_eval_batch launch equals
    the point-by-point evaluation and the oracle's driver."""
    from qinchworm_b200.inchworm import Solver, correlator_2p, inchworm
    ex, grid, f = models.bethe_two_state(n_tau=24)
    solver = Solver(ex, ctx=gpu_ctx)
    inchworm(ex, grid, range(0, 4), range(0, 4), 2 ** 8, solver=solver)
    g_b = correlator_2p(ex, grid, range(0, 4), 2 ** 8, solver=solver)[0]
    g_s = correlator_2p(ex, grid, range(0, 4), 2 ** 8, solver=solver, batch=False)[0]
    assert relerr(g_b, g_s) < 1e-13
    ref = oracle_lib.correlator_2p(ex.flatten(), ex.P, range(0, 4), 2 ** 8)
    assert relerr(g_b, ref) < RTOL


def test_edge_cases(gpu_ctx, qlib, oracle_lib, monkeypatch):
    """Ragged and degenerate inputs: a single sample, an empty sample range, order-0-only calls, the
    first Sobol point (x = 0 collapses every time onto its lower bound), a scrambled sequence in the
    device-resident run, the complex block kernel, and argument errors."""
    from qinchworm_b200.inchworm import Solver
    ex, grid, f = models.anderson(n_tau=16)
    pl = gpu_ctx.set_expansion(ex)
    o = oracle_lib.Oracle(pl, ex.P)
    ids = []
    for order in range(0, 3):
        for k in ([0] if order == 0 else range(1, 2 * order)):
            pr, pa = qlib.topologies(order, k)
            gpu_ctx.set_topologies(len(ids), qlib.MODE_BOLD, order, k, pr, pa)
            o.set_topologies(len(ids), qlib.MODE_BOLD, order, k, pr, pa)
            ids.append(len(ids))
    tau = grid.tau
    # N = 1: only the first point of the sequence (all times on their lower bounds)
    assert relerr(gpu_ctx.eval(0.0, tau[7], tau[8], ids, 1), o.eval(0.0, tau[7], tau[8], ids, 1)) < RTOL
    # empty range: sampled entries contribute nothing, the exact order-0 entry is still evaluated
    got = gpu_ctx.eval_range(0.0, tau[7], tau[8], ids, 64, 10, 0)
    ref = o.eval(0.0, tau[7], tau[8], ids, 64, start=10, count=0)
    assert np.abs(got - ref).max() < 1e-300 + RTOL * np.abs(ref).max()
    # order 0 alone
    assert relerr(gpu_ctx.eval(0.0, tau[7], tau[8], ids[:1], 8), o.eval(0.0, tau[7], tau[8], ids[:1], 8)) < RTOL
    # a non-power-of-two range that is not a multiple of the 32-sample blocks, at the end of the sequence
    got = gpu_ctx.eval_range(0.0, tau[3], tau[4], ids, 1000, 1000 - 77, 77)
    ref = o.eval(0.0, tau[3], tau[4], ids, 1000, start=1000 - 77, count=77)
    assert relerr(got, ref) < RTOL
    # batched evaluation with a scrambled sequence
    rng = np.random.default_rng(5)
    sob = []
    for e in ids:
        D = 2 * gpu_ctx.entry_order[e]
        m = qlib.sobol_direction_numbers(D)
        sob.append(qlib.sobol_scramble(m, rng.integers(0, 2, (D, 32)), rng.integers(0, 2, (D, 32, 32))) if D
                   else (m, np.zeros(0, dtype=np.uint32)))
    times = np.array([[0.0, tau[k], tau[k + 1]] for k in (2, 9, 13)])
    got = gpu_ctx.eval_batch(times, ids, 128, sobol=sob)
    for z in range(3):
        ref = o.eval(times[z, 0], times[z, 1], times[z, 2], ids, 128, sobol=sob)
        assert relerr(got[z], ref) < RTOL
    # argument errors are reported, not crashes
    with pytest.raises(qlib.QiwError):
        gpu_ctx.eval(0.0, tau[7], tau[8], [999], 8)
    with pytest.raises(qlib.QiwError):
        gpu_ctx.eval(0.0, tau[7], tau[8], [0, 0], 8)
    with pytest.raises(qlib.QiwError):
        gpu_ctx.eval_range(0.0, tau[7], tau[8], ids, 8, 5, 9)
    # block model through the general complex kernel (the real-arithmetic walker is the default)
    ex2, grid2, _ = models.two_level_mixed(n_tau=10, theta=0.4)
    pl2 = gpu_ctx.set_expansion(ex2)
    o2 = oracle_lib.Oracle(pl2, ex2.P)
    ids2 = []
    for order in range(0, 3):
        for k in ([0] if order == 0 else range(1, 2 * order)):
            pr, pa = qlib.topologies(order, k)
            gpu_ctx.set_topologies(len(ids2), qlib.MODE_BOLD, order, k, pr, pa)
            o2.set_topologies(len(ids2), qlib.MODE_BOLD, order, k, pr, pa)
            ids2.append(len(ids2))
    ref = o2.eval(0.0, grid2.tau[4], grid2.tau[5], ids2, 64)
    got_real = gpu_ctx.eval(0.0, grid2.tau[4], grid2.tau[5], ids2, 64)
    monkeypatch.setenv("QIW_FORCE_COMPLEX", "1")
    got_cplx = gpu_ctx.eval(0.0, grid2.tau[4], grid2.tau[5], ids2, 64)
    monkeypatch.delenv("QIW_FORCE_COMPLEX")
    assert relerr(got_real, ref) < RTOL and relerr(got_cplx, ref) < RTOL


def test_full_size_properties(gpu_ctx, qlib):
    """BASELINE.json configs[0] at full size (README Anderson model, n_tau = 200, orders 0:4, N = 2^10), where
    the oracle would take minutes: size-independent properties instead.
      * Z and rho_imp against the regression anchors of SURVEY.md (R3: restatement predictions for the
        current algorithm, good to ~1e-6) and rho against README.md:204 (2e-5);
      * Tr rho = 1, spin symmetry rho_up = rho_dn exactly (the two flavours run through identical arithmetic);
      * bit-identical results run to run (fixed-order reductions);
      * additivity over disjoint Sobol index ranges at N = 2^17 (what the multi-GPU sharding relies on)."""
    from qinchworm_b200 import ppgf
    from qinchworm_b200.inchworm import Solver, inchworm, _bold_entries
    ex, grid, f = models.anderson(n_tau=200)
    P0 = ex.P.copy()
    solver = Solver(ex, ctx=gpu_ctx)
    inchworm(ex, grid, range(0, 5), range(0, 5), 2 ** 10, solver=solver)
    P1 = ex.P.copy()
    Z = ppgf.partition_function(ex)
    assert abs(Z - 1.9428382858) < 5e-6 * 1.94
    ppgf.normalize(ex)
    rho = np.array([d[0, 0].real for d in ppgf.density_matrix(ex)])
    assert np.abs(rho - np.array([0.5147108781, 0.2304382136, 0.2304382136, 0.0244126947])).max() < 2e-6
    assert np.abs(rho - np.array([0.5147267752890132, 0.23043341978635334, 0.23043341978635334, 0.02440638513828009])).max() < 3e-5
    assert abs(rho.sum() - 1.0) < 1e-12 and rho[1] == rho[2]
    ex.P[:] = P0
    inchworm(ex, grid, range(0, 5), range(0, 5), 2 ** 10, solver=solver)
    assert np.array_equal(ex.P, P1)
    # additivity over sample ranges with the machine full
    ex.P[:] = P1
    solver.upload_P()
    N = 2 ** 17
    bold = _bold_entries(solver, range(0, 5), N, None, None)
    ids = [t.entry_id for t in bold]
    tau = grid.tau
    whole = gpu_ctx.eval(0.0, tau[120], tau[121], ids, N)
    cuts = [0, 12345, 65536, 100001, N]
    parts = sum(gpu_ctx.eval_range(0.0, tau[120], tau[121], ids, N, a, b - a) for a, b in zip(cuts[:-1], cuts[1:]))
    parts[0] = whole[0]      # the exact order-0 entry is evaluated in full by every range call
    assert relerr(parts, whole) < 1e-12


def test_randomised_sequences_in_one_launch(gpu_ctx, qlib, oracle_lib):
    """qiw_eval_seqs (all scrambled sequences of mean_std_from_randomization in one launch) equals
    one qiw_eval per sequence and the oracle; the host driver returns mean and std over the sequences."""
    from qinchworm_b200.inchworm import RandomizationParams, Solver, inchworm_step, _bold_entries
    ex, grid, f = models.anderson(n_tau=20)
    solver = Solver(ex, ctx=gpu_ctx)
    o = oracle_lib.Oracle(solver.payload, ex.P)
    rp = RandomizationParams(rng=np.random.default_rng(3), N_seqs=5, target_std=0.0)
    top_data = _bold_entries(solver, range(0, 4), 2 ** 8, rp, None)
    ids = [td.entry_id for td in top_data]
    for j, td in enumerate(top_data):
        o.set_topologies(j, qlib.MODE_BOLD, td.order, td.n_pts_after, td.topologies[0], td.topologies[1])
    rng = np.random.default_rng(17)
    seqs = []
    for s_ in range(3):
        one = []
        for td in top_data:
            D = 2 * td.order
            m = qlib.sobol_direction_numbers(D)
            one.append(qlib.sobol_scramble(m, rng.integers(0, 2, (D, 32)), rng.integers(0, 2, (D, 32, 32))) if D
                       else (m, np.zeros(0, dtype=np.uint32)))
        seqs.append(one)
    tau = grid.tau
    got = gpu_ctx.eval_seqs(0.0, tau[9], tau[10], ids, 2 ** 8, seqs)
    for z in range(3):
        one = gpu_ctx.eval(0.0, tau[9], tau[10], ids, 2 ** 8, sobol=seqs[z])
        ref = o.eval(0.0, tau[9], tau[10], list(range(len(ids))), 2 ** 8, sobol=seqs[z])
        assert relerr(got[z], one) < 1e-13 and relerr(got[z], ref) < RTOL
    # afterwards the default (unscrambled) sequence must be back in place
    assert relerr(gpu_ctx.eval(0.0, tau[9], tau[10], ids, 2 ** 8), o.eval(0.0, tau[9], tau[10], list(range(len(ids))), 2 ** 8)) < RTOL
    value, contribs, contribs_std = inchworm_step(solver, grid, 0, 9, 10, top_data)
    assert np.isfinite(contribs_std[3]).all() and np.abs(contribs_std[3]).max() > 0 and np.abs(contribs_std[0]).max() == 0


@pytest.mark.parametrize("arith", ["real", "complex"])
def test_paired_records(gpu_ctx, qlib, oracle_lib, monkeypatch, arith):
    """Records shared by two initial sectors (one set of pair-interaction operands for both; default from order 5 on)
    forced on for every order, in both arithmetic modes, bold + bare + correlator, plus one genuine order-5 entry."""
    monkeypatch.setenv("QIW_LANE_DUAL", "1")
    if arith == "complex":
        monkeypatch.setenv("QIW_FORCE_COMPLEX", "1")
    ex, grid, f = models.anderson(n_tau=30, corr=True)
    rng = np.random.default_rng(21)
    ex.P = ex.P * (1.0 + 0.05 * rng.random(ex.P.shape))
    pl = gpu_ctx.set_expansion(ex)
    o = oracle_lib.Oracle(pl, ex.P)
    tau = grid.tau
    eid = 0
    for mode, (ki, kw, kf), orders in ((qlib.MODE_BOLD, (0, 14, 15), range(0, 5)), (qlib.MODE_BARE, (0, 0, 1), range(0, 4)),
                                       (qlib.MODE_CORR, (0, 11, len(tau) - 1), range(0, 4)), (qlib.MODE_BOLD, (0, 20, 21), [5])):
        ids = []
        for order in orders:
            ks = [None] if mode == qlib.MODE_BARE else ([0] if order == 0 else ([3] if order == 5 else range(1, 2 * order)))
            for k in ks:
                pr, pa = qlib.topologies(order, None if mode == qlib.MODE_BARE else k, mode == qlib.MODE_CORR)
                if len(pa) == 0:
                    continue
                kk = 2 * order if mode == qlib.MODE_BARE else k
                gpu_ctx.set_topologies(eid, mode, order, kk, pr, pa)
                o.set_topologies(eid, mode, order, kk, pr, pa)
                ids.append(eid)
                eid += 1
        N = 2 ** 7 if 5 in orders else 2 ** 9
        got = gpu_ctx.eval(tau[ki], tau[kw], tau[kf], ids, N)
        ref = o.eval(tau[ki], tau[kw], tau[kf], ids, N)
        assert relerr(got, ref) < RTOL, (mode, relerr(got, ref))
    lp = gpu_ctx.entry_lane_program(ids[-1])
    assert len(lp["sections"]) > 0 and any(M == 4 for _, M, _, _ in lp["sections"])
    assert any(int(sc) >> 16 and M == 4 for sc, M, _, _ in lp["sections"]) and any(int(sc) >> 16 and M == 2 for sc, M, _, _ in lp["sections"])


@pytest.mark.parametrize("spline", [True, False])
def test_two_site_dimer_physics(gpu_ctx, qlib, oracle_lib, spline):
    """test/dimers.jl:34-119 through the product path: rho of the two-site dimer (orders 0:3, n_tau = 32, N = 128,
    spline-interpolated and plain grid functions) within the reference's 1e-4 of exact diagonalisation, and P(tau)
    equal to the oracle's."""
    from qinchworm_b200 import ppgf
    from qinchworm_b200.inchworm import Solver, inchworm
    ex, grid, f = models.single_level(n_tau=32, beta=1.0, mu=-0.5, eps=2.0, V=0.5, spline=spline, rev="reverse")
    ref = oracle_lib.inchworm(ex.flatten(), ex.P, range(0, 4), range(0, 4), 8 * 2 ** 4)["P"]
    inchworm(ex, grid, range(0, 4), range(0, 4), 8 * 2 ** 4, solver=Solver(ex, ctx=gpu_ctx))
    assert relerr(ex.P, ref) < RTOL
    ppgf.normalize(ex)
    rho = np.array([d[0, 0].real for d in ppgf.density_matrix(ex)])
    assert np.abs(rho - models.two_site_dimer_exact_rho()).max() < 1e-4


@pytest.mark.parametrize("arith", ["real", "complex"])
@pytest.mark.parametrize("order", [5, 6])
def test_high_order_entries_vs_oracle(gpu_ctx, qlib, oracle_lib, monkeypatch, arith, order):
    """BASELINE.json configs[4] (stress, orders up to 6): EVERY bold entry of order 5 (9 entries) and order 6
    (11 entries) against the oracle at N = 2^6, entry by entry and element-wise, in real and in complex
    arithmetic.  These orders run through their own record format / kernel instantiation."""
    import os
    if arith == "complex":
        monkeypatch.setenv("QIW_FORCE_COMPLEX", "1")
    ex, grid, f = models.anderson(n_tau=40)
    rng = np.random.default_rng(50 + order)
    ex.P = ex.P * (1.0 + 0.05 * rng.random(ex.P.shape))
    pl = gpu_ctx.set_expansion(ex)
    o = oracle_lib.Oracle(pl, ex.P, threads=os.cpu_count() or 1)
    ids = []
    for k in range(1, 2 * order):
        pr, pa = qlib.topologies(order, k)
        assert len(pa) > 0
        gpu_ctx.set_topologies(len(ids), qlib.MODE_BOLD, order, k, pr, pa)
        o.set_topologies(len(ids), qlib.MODE_BOLD, order, k, pr, pa)
        ids.append(len(ids))
    assert len(ids) == 2 * order - 1
    tau = grid.tau
    N = 2 ** 6
    got = gpu_ctx.eval(0.0, tau[23], tau[24], ids, N)
    ref = o.eval(0.0, tau[23], tau[24], ids, N)
    for j in ids:
        assert relerr_elem(got[j], ref[j]) < RTOL, (order, j + 1, relerr_elem(got[j], ref[j]))


def test_c1_full_size_vs_oracle(gpu_ctx, qlib, oracle_lib):
    """BASELINE.json configs[0] at FULL size (README.md:34-158: n_tau = 200, orders 0:4, N = 2^10): P(tau) for
    all 200 grid points and 4 sectors, Z, rho_imp and G(tau) (correlator_2p, orders 0:3) of the CUDA path
    against the oracle's own drivers, element-wise, tolerance 1e-10."""
    import os
    from qinchworm_b200 import ppgf
    from qinchworm_b200.expansion import add_corr_operators
    from qinchworm_b200.inchworm import Solver, correlator_2p, inchworm
    threads = os.cpu_count() or 1
    ex, grid, f = models.anderson(n_tau=200)
    ref = oracle_lib.inchworm(ex.flatten(), ex.P, range(0, 5), range(0, 5), 2 ** 10, threads=threads)["P"]
    solver = Solver(ex, ctx=gpu_ctx)
    inchworm(ex, grid, range(0, 5), range(0, 5), 2 ** 10, solver=solver)
    assert relerr_elem(ex.P, ref) < RTOL, relerr_elem(ex.P, ref)
    Z, Zref = ppgf.partition_function(ex), (1j * ref[-1]).sum()
    assert abs(Z - Zref) < RTOL * abs(Zref)
    rho = np.array([d[0, 0] for d in ppgf.density_matrix(ex)]) / Z     # rho_imp = i P(beta) / Z (README.md:150-158)
    rho_ref = 1j * ref[-1] / Zref
    assert relerr_elem(rho, rho_ref) < RTOL
    # G(tau) on the converged P, both sides starting from the SAME table (the oracle's)
    ex.P[:] = ref
    add_corr_operators(ex, (f.c("up"), f.c_dag("up")))
    g = correlator_2p(ex, grid, range(0, 4), 2 ** 10, solver=solver)[0]
    g_ref = oracle_lib.correlator_2p(ex.flatten(), ref, range(0, 4), 2 ** 10, threads=threads)
    assert relerr_elem(g, g_ref) < RTOL, relerr_elem(g, g_ref)


def test_c4_order4_bold_step_vs_oracle(gpu_ctx, qlib, oracle_lib):
    """BASELINE.json configs[3] (two-band e_g model, 9 sectors with blocks 1/2/4, 16 pairs) at ORDER 4: one bold step
    with all seven order-4 entries (21.3 M configurations per sample-set) against the oracle at N = 2^4."""
    import os
    ex, grid, f = models.two_band(n_tau=10)
    rng = np.random.default_rng(44)
    ex.P = ex.P * (1.0 + 0.05 * rng.random(ex.P.shape))
    pl = gpu_ctx.set_expansion(ex)
    o = oracle_lib.Oracle(pl, ex.P, threads=os.cpu_count() or 1)
    ids = []
    for k in range(1, 8):
        pr, pa = qlib.topologies(4, k)
        gpu_ctx.set_topologies(len(ids), qlib.MODE_BOLD, 4, k, pr, pa)
        o.set_topologies(len(ids), qlib.MODE_BOLD, 4, k, pr, pa)
        ids.append(len(ids))
    tau = grid.tau
    N = 2 ** 4
    got = gpu_ctx.eval(0.0, tau[5], tau[6], ids, N)
    ref = o.eval(0.0, tau[5], tau[6], ids, N)
    for j in ids:
        assert relerr(got[j], ref[j]) < RTOL, (j + 1, relerr(got[j], ref[j]))
    tot, tot_ref = got.sum(axis=0), ref.sum(axis=0)
    assert relerr_elem(tot, tot_ref) < RTOL


@pytest.mark.gpu
@pytest.mark.parametrize("arith", ["real", "complex"])
def test_run_kernel_vs_step_launches(gpu_ctx, qlib, oracle_lib, monkeypatch, arith):
    """qiw_inchworm_run: the persistent cooperative run kernel (all bold steps in one launch; P and the pair-interaction
    tables staged in shared memory, one grid barrier per step) against one step kernel per step (QIW_NO_RUN_KERNEL=1)
    and against the oracle, in both arithmetic modes; sample counts with one job per CTA and with several; spline-interpolated
    pair-interaction tables (staged in shared memory as values + second derivatives)."""
    from qinchworm_b200.inchworm import MODE_BARE, Solver, _bold_entries
    if arith == "complex":
        monkeypatch.setenv("QIW_FORCE_COMPLEX", "1")
    ex, grid, f = models.anderson(n_tau=40)
    P0 = ex.P.copy()
    solver = Solver(ex, ctx=gpu_ctx)
    for N in (2 ** 8, 2 ** 12):
        bare = [solver.make_entry(MODE_BARE, o, 2 * o, N) for o in range(4)]
        bold = _bold_entries(solver, range(4), N, None, None)
        res = {}
        for mode in ("run", "step"):
            monkeypatch.setenv("QIW_NO_RUN_KERNEL", "0" if mode == "run" else "1")
            gpu_ctx.set_P(0, P0)
            l0 = gpu_ctx.launch_count()
            hist = gpu_ctx.inchworm_run([t.entry_id for t in bare], [t.entry_id for t in bold], N, want_contribs=True)
            res[mode] = (gpu_ctx.get_P(), hist, gpu_ctx.launch_count() - l0)
        assert res["run"][2] == 2 and res["step"][2] == grid.n_tau - 1       # bare step + one run kernel
        assert relerr(res["run"][0], res["step"][0]) < 1e-13
        assert relerr(res["run"][1], res["step"][1]) < 1e-13
        if N == 2 ** 8:
            ref = oracle_lib.inchworm(ex.flatten(), P0, range(4), range(4), N)["P"]
            assert relerr(res["run"][0], ref) < RTOL
    # spline-interpolated pair interactions (the reference's golden configuration): the splines' values and second
    # derivatives are staged in shared memory like the grid tables; run kernel vs per-step launches vs test/inchworm.h5
    ex2, grid2, _ = models.single_level(n_tau=20, spline=True)
    solver2 = Solver(ex2, ctx=gpu_ctx)
    bare = [solver2.make_entry(MODE_BARE, o, 2 * o, 2 ** 8) for o in range(3)]
    bold = _bold_entries(solver2, range(4), 2 ** 8, None, None)
    G = load_golden("inchworm_h5.json")
    res2 = {}
    for mode in ("run", "step"):
        monkeypatch.setenv("QIW_NO_RUN_KERNEL", "0" if mode == "run" else "1")
        solver2.upload_P()
        l0 = gpu_ctx.launch_count()
        gpu_ctx.inchworm_run([t.entry_id for t in bare], [t.entry_id for t in bold], 2 ** 8, want_contribs=False)
        assert gpu_ctx.launch_count() - l0 == (2 if mode == "run" else grid2.n_tau - 1)
        res2[mode] = gpu_ctx.get_P()
        assert relerr(res2[mode][:, 1], G["/inchworm/P/1"].ravel()) < RTOL
    assert relerr(res2["run"], res2["step"]) < 1e-13


@pytest.mark.gpu
def test_block_model_single_launch_and_batching(gpu_ctx, qlib, oracle_lib):
    """Sector blocks larger than 1x1 (C4-type models): one launch per step (reduction, all-reduce and P update fused into
    block_walk_kernel's tail), and the batched forms — all grid points of a correlator / all scrambled sequences in one
    launch (gridDim.z) — equal to the one-by-one evaluation and to the oracle."""
    from qinchworm_b200.inchworm import Solver, correlator_2p, inchworm
    ex, grid, f = models.two_band(n_tau=10)
    solver = Solver(ex, ctx=gpu_ctx)
    l0 = gpu_ctx.launch_count()
    inchworm(ex, grid, range(0, 3), range(0, 3), 2 ** 7, solver=solver, device_resident=True)
    assert gpu_ctx.launch_count() - l0 == grid.n_tau - 1                      # one kernel per step, nothing else
    refP = oracle_lib.inchworm(ex.flatten(), models.two_band(n_tau=10)[0].P, range(0, 3), range(0, 3), 2 ** 7)["P"]
    assert relerr(ex.P, refP) < RTOL
    l0 = gpu_ctx.launch_count()
    g_b = correlator_2p(ex, grid, range(0, 3), 2 ** 7, solver=solver)[0]
    n_batched = gpu_ctx.launch_count() - l0
    g_s = correlator_2p(ex, grid, range(0, 3), 2 ** 7, solver=solver, batch=False)[0]
    assert n_batched <= 2                                                     # tau = 0 (order 0 only) + all other grid points
    assert relerr(g_b, g_s) < 1e-13
    assert relerr(g_b, oracle_lib.correlator_2p(ex.flatten(), ex.P, range(0, 3), 2 ** 7)) < RTOL
    # scrambled sequences: qiw_eval_seqs in one launch vs one qiw_eval per sequence
    ids, tds = [], []
    for order in range(0, 3):
        for k in ([0] if order == 0 else range(1, 2 * order)):
            td = solver.make_entry(qlib.MODE_BOLD, order, k, 2 ** 7)
            if td is not None:
                ids.append(td.entry_id); tds.append(td)
    rng = np.random.default_rng(5)
    seqs = []
    for s_ in range(3):
        one = []
        for td in tds:
            D = 2 * td.order
            m = qlib.sobol_direction_numbers(D)
            one.append(qlib.sobol_scramble(m, rng.integers(0, 2, (D, 32)), rng.integers(0, 2, (D, 32, 32))) if D
                       else (m, np.zeros(0, dtype=np.uint32)))
        seqs.append(one)
    tau = grid.tau
    l0 = gpu_ctx.launch_count()
    got = gpu_ctx.eval_seqs(0.0, tau[5], tau[6], ids, 2 ** 7, seqs)
    assert gpu_ctx.launch_count() - l0 == 1
    for z in range(3):
        assert relerr(got[z], gpu_ctx.eval(0.0, tau[5], tau[6], ids, 2 ** 7, sobol=seqs[z])) < 1e-13


@pytest.mark.gpu
def test_step_seam_scale_P(gpu_ctx, qlib):
    """qiw_scale_P (set_ppgf! + normalize! across the step seam): row k_f := row, then every stored row k times
    exp(-lambda tau_k) — the device's table follows the host's without a re-upload; N_samples = 0 evaluates order 0 only,
    as the reference does (src/inchworm.jl:159,258)."""
    from qinchworm_b200.inchworm import Solver, inchworm
    ex, grid, f = models.anderson(n_tau=16)
    solver = Solver(ex, ctx=gpu_ctx)
    solver.upload_P()
    P = ex.P.copy()
    row = (0.3 + 0.1 * np.arange(gpu_ctx.bsize)) * 1j
    lam = 0.37
    gpu_ctx.scale_P(5, row, lam)
    P[5] = row
    P *= np.exp(-grid.tau * lam)[:, None]
    assert relerr(gpu_ctx.get_P(), P) < 1e-15
    gpu_ctx.scale_P(7, row)                       # lambda = 0: the row alone
    P[7] = row
    assert relerr(gpu_ctx.get_P(), P) < 1e-15
    ex0, grid0, _ = models.anderson(n_tau=16)
    ex1, grid1, _ = models.anderson(n_tau=16)
    Po0, _ = inchworm(ex0, grid0, range(0, 4), range(0, 4), 0, solver=Solver(ex0, ctx=gpu_ctx))
    Po1, _ = inchworm(ex1, grid1, [0], [0], 2 ** 6, solver=Solver(ex1, ctx=gpu_ctx))
    assert relerr(ex0.P, ex1.P) < 1e-15 and set(Po0) == {0}


@pytest.mark.gpu
def test_randomisation_early_stop_per_entry(gpu_ctx, qlib):
    """mean_std_from_randomization is called per entry (src/inchworm.jl:174, src/randomization.jl:93-99): with a target_std
    every entry stops on its own criterion — here a huge target stops every sampled entry after its second sequence, so the
    number of library calls is 1 (order 0, exact, no sequence drawn) + 2 per sampled entry — and a host RNG is consumed
    entry by entry."""
    from qinchworm_b200.inchworm import RandomizationParams, Solver, _bold_entries, _scrambled_sequence
    ex, grid, f = models.anderson(n_tau=16)
    solver = Solver(ex, ctx=gpu_ctx)
    rp = RandomizationParams(rng=np.random.default_rng(11), N_seqs=6, target_std=1e6)
    top = _bold_entries(solver, range(0, 3), 2 ** 6, rp, None)
    l0 = gpu_ctx.launch_count()
    smp = solver.eval_samples(0.0, grid.tau[7], grid.tau[8], top)
    assert gpu_ctx.launch_count() - l0 == 1 + 2 * (len(top) - 1)
    assert [len(x) for x in smp] == [1] + [2] * (len(top) - 1)
    # the same stream drawn by hand in the reference's order: entry by entry, two sequences each
    rng = np.random.default_rng(11)
    for td, x in zip(top[1:], smp[1:]):
        for s_ in range(2):
            seq = _scrambled_sequence(2 * td.order, rng)
            one = gpu_ctx.eval(0.0, grid.tau[7], grid.tau[8], [td.entry_id], 2 ** 6, sobol=[seq])[0]
            assert relerr(x[s_], one) < 1e-15
    mean, std = solver.eval_entries(0.0, grid.tau[7], grid.tau[8], top)
    assert np.abs(std[0]).max() == 0 and np.isfinite(std[1:]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("arith", ["real", "complex"])
def test_blocks_larger_than_4x4(gpu_ctx, qlib, oracle_lib, monkeypatch, arith):
    """Real arithmetic: block_mma_kernel (FP64 tensor cores, mma.sync.m8n8k4.f64, warp = sample); complex arithmetic
    (QIW_FORCE_COMPLEX=1): the general block_step_kernel<8>.
    Sector blocks of 5 to 8 rows (north_star's "sector blocks large enough to be a dense contraction"): the two-band
    model resolved by particle number only has blocks {1,4,6,4,1}.  Bold / bare / correlator steps against the oracle
    (itself checked against the brute-force Fock-space formula for this model, tests/test_oracle_golden.py), and — a
    size-independent property — the whole run must give the same partition function and the same density matrix in the
    Fock basis as the 9-sector bookkeeping {1,1,1,1,2,2,2,2,4} of the same model on the same Sobol points."""
    from qinchworm_b200 import ppgf
    from qinchworm_b200.inchworm import Solver, inchworm
    if arith == "complex":
        monkeypatch.setenv("QIW_FORCE_COMPLEX", "1")
    ex, grid, f = models.two_band(n_tau=10, big_blocks=True)
    assert max(ex.dims) == 6
    solver = Solver(ex, ctx=gpu_ctx)
    o = oracle_lib.Oracle(solver.payload, ex.P)
    tau = grid.tau
    eid = 0
    for mode, (ki, kw, kf) in ((qlib.MODE_BOLD, (0, 5, 6)), (qlib.MODE_BARE, (0, 0, 1)), (qlib.MODE_CORR, (0, 4, 9))):
        ids = []
        for order in range(0, 3):
            ks = [None] if mode == qlib.MODE_BARE else ([0] if order == 0 else (1, 2 * order - 1))
            for k in ks:
                pr, pa = qlib.topologies(order, None if mode == qlib.MODE_BARE else k, mode == qlib.MODE_CORR)
                if len(pa) == 0:
                    continue
                kk = 2 * order if mode == qlib.MODE_BARE else k
                gpu_ctx.set_topologies(200 + eid, mode, order, kk, pr, pa)
                o.set_topologies(eid, mode, order, kk, pr, pa)
                ids.append(eid)
                eid += 1
        got = gpu_ctx.eval(tau[ki], tau[kw], tau[kf], [200 + i for i in ids], 2 ** 6)
        ref = o.eval(tau[ki], tau[kw], tau[kf], ids, 2 ** 6)
        assert relerr(got, ref) < RTOL, (mode, relerr(got, ref))
        prof = gpu_ctx.profile_read(reset=True) if False else None
    rho = []
    for big in (True, False):
        exr, gridr, _ = models.two_band(n_tau=8, big_blocks=big)
        inchworm(exr, gridr, range(0, 3), range(0, 3), 2 ** 6, solver=Solver(exr, ctx=gpu_ctx))
        Z = ppgf.partition_function(exr)
        rho.append(exr.ed.to_fock_basis(ppgf.density_matrix(exr)) / Z)
        assert abs(np.trace(rho[-1]) - 1.0) < 1e-13
    assert np.abs(rho[0] - rho[1]).max() < 1e-10
