"""C3 (bench/fermi_hubbard_dimer) and C4 (bench/two_band_eg_model_discrete_bath) as whole inchworm! runs:
device-resident run on one GPU next to the CPU oracle port on all host cores (bounded number of bold steps).

usage: bench_c34.py c3|c4 [max_order] [n_tau] [N_samples] [cpu_bold_steps]
Prints one JSON line per configuration (kept under profiles/ as r1_bench_c3.json / r1_bench_c4.json)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import models
from oracle import oracle as orc
from qinchworm_b200 import lib, ppgf
from qinchworm_b200.inchworm import MODE_BARE, RandomizationParams, Solver, _bold_entries, inchworm

which = sys.argv[1] if len(sys.argv) > 1 else "c4"
max_order = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n_tau = int(sys.argv[3]) if len(sys.argv) > 3 else (64 if which == "c3" else 32)
N = int(sys.argv[4]) if len(sys.argv) > 4 else (2 ** 15 if which == "c3" else 2 ** 10)
cpu_steps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
make = (lambda: models.hubbard_dimer_impurity(n_tau=n_tau)) if which == "c3" else (lambda: models.two_band(n_tau=n_tau))
orders = range(0, max_order + 1)

ex, grid, f = make()
ctx = lib.Context(device=0)
peak = ctx.measure_fp64_peak()
solver = Solver(ex, ctx=ctx)
t0 = time.perf_counter()
rp = RandomizationParams()
bare = [solver.make_entry(MODE_BARE, o, 2 * o, N, rp) for o in orders]
bold = _bold_entries(solver, orders, N, rp, None)
t_compile = time.perf_counter() - t0
st_bare = [ctx.entry_stats(t.entry_id) for t in bare]
st_bold = [ctx.entry_stats(t.entry_id) for t in bold]
n_of = lambda td: N if td.order > 0 else 1
evals_bare = sum(n_of(t) * s["n_top"] for t, s in zip(bare, st_bare))
evals_bold = sum(n_of(t) * s["n_top"] for t, s in zip(bold, st_bold))
flops_bold = sum(n_of(t) * s["flops_per_sample"] for t, s in zip(bold, st_bold))
evals_run = evals_bare + (n_tau - 2) * evals_bold
P0 = ex.P.copy()
ms = []
for rep in range(3):
    ex.P[:] = P0
    t = time.perf_counter()
    inchworm(ex, grid, orders, orders, N, solver=solver, device_resident=True)
    ms.append((time.perf_counter() - t) * 1e3)
wall = float(np.median(ms[1:]))
dev = ctx.last_device_ms()
out = {"config": which, "orders": "0:%d" % max_order, "n_tau": n_tau, "N_samples": N, "sector_dims": sorted(int(d) for d in ex.dims),
       "gpu": {"inchworm_wall_ms": wall, "device_ms": dev, "diagram_evals_per_s": evals_run / (wall * 1e-3),
               "algorithmic_tflops_bold": flops_bold * (n_tau - 2) / (dev * 1e-3) / 1e12,
               "frac_of_measured_fp64_peak": flops_bold * (n_tau - 2) / (dev * 1e-3) / 1e12 / peak, "fp64_peak_tflops": peak,
               "compile_s": t_compile, "launches": ctx.launch_count()}}
if which == "c3":
    ppgf.normalize(ex)
    rho = ex.ed.to_fock_basis(ppgf.density_matrix(ex))
    out["max_abs_rho_diff_vs_exact_ed"] = float(np.abs(rho - models.hubbard_dimer_exact_rho()).max())
else:
    ppgf.normalize(ex)
    out["trace_rho"] = float(sum(np.trace(d).real for d in ppgf.density_matrix(ex)))
# CPU oracle port: bare step + the first cpu_steps bold steps on all host cores
ex2 = make()[0]
cores = os.cpu_count() or 1
t = time.perf_counter()
r = orc.inchworm(ex2.flatten(), ex2.P, orders, orders, N, threads=cores, max_bold_steps=cpu_steps)
dt = time.perf_counter() - t
out["cpu_port"] = {"cores": cores, "seconds": dt, "diagram_evals_per_s": r["evals"] / dt,
                   "sample": "bare step + first %d of %d bold steps" % (cpu_steps, n_tau - 2)}
out["gpu_over_cpu_port"] = out["gpu"]["diagram_evals_per_s"] / out["cpu_port"]["diagram_evals_per_s"]
# parity of the steps both sides computed (same inputs): rows 0 .. cpu_steps + 1 of the P table
ex3 = make()[0]
rows = cpu_steps + 2
if rows <= n_tau:
    ex3.P[:] = P0
    Po, _ = inchworm(ex3, grid, orders, orders, N, solver=Solver(ex3, ctx=ctx), device_resident=True)
    gpu_rows = sum(Po.values())[:rows]
    cpu_rows = sum(r["P_orders"].values())[:rows]
    out["max_rel_diff_first_rows"] = float(np.abs(gpu_rows[1:] - cpu_rows[1:]).max() / np.abs(cpu_rows[1:]).max())
print(json.dumps(out))
