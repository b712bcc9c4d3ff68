#!/usr/bin/env python
"""bench.py — qMC diagram evaluations/sec and inchworm! wall time on the README Anderson configuration.

Workload (BASELINE.json configs[0], the configuration the metric is quoted on): single-orbital
Anderson model on a Bethe bath, beta=10, U=1, eps=0.1, V=0.5, n_tau=200, orders 0:4, orders_bare 0:4,
N_samples = 2^10 per GPU (weak scaling: N_samples = 2^10 * n_gpus, each GPU takes a disjoint Sobol
index range, one ncclAllReduce per inchworm step).  One "step" = one complete inchworm! run
(bare step + 198 bold steps) = 5.69e7 diagram evaluations per 2^10 samples.

    value : diagram evaluations/s of the device-resident run (qiw_inchworm_run; tables resident in
            HBM, CUDA events on the library's stream around the whole run), max over ranks.  The run is two launches: one
            step kernel for the bare step, one persistent cooperative run kernel for the 198 bold steps
    e2e   : the same metric through the public host API inchworm(expansion, grid, orders, orders_bare,
            N_samples) with HOST buffers: the atomic P table goes host->device, the final P table and the
            order-resolved contributions come back device->host inside the timed region (compiled
            entries are cached in the Solver, like the context).  `e2e_stepped` is the same call with
            device_resident=False: one qiw_eval per step through the three-worker seam, set_ppgf!/normalize! on the host,
            the device's table following with qiw_scale_P (one row + lambda) — what the thin Julia shim does.
    roofline     : dominant kernel (the run kernel) against the FP64 FMA peak measured in the same process by a
                   DFMA-saturating kernel: `frac` credits the reference's complex chain, `executed` is what the kernel
                   executes (FP64 instructions, shared-memory operands) next to ncu's counters of the committed capture
    cpu_baseline : the CPU oracle port (faithful restatement of the reference algorithm), all host
                   cores, on a bounded sample of the same workload

    parity       : the same workload on the CPU oracle (the cpu_baseline leg runs it anyway): P(tau), Z, rho_imp and
                   G(tau) element-wise against the GPU's, tolerance 1e-10; above it the bench exits non-zero.  At N > 1
                   GPUs rank 0 checks the first bold steps of the sharded run against the oracle at the same N_samples.
    stress_c5_step : BASELINE.json configs[4] (orders 0:6, n_tau = 400, N_samples = 2^20 FIXED, i.e. strong scaling over
                   the GPUs of the job), a few bold steps in the middle of the run, with its own oracle check at N = 2^6.
    other_configs : (one GPU) BASELINE.json configs[1..3] as short whole-workload runs with their own checks: C2 batched
                   correlator_2p sweep against the oracle, C3 Hubbard dimer orders 0:4 against exact diagonalisation, C4
                   two-band model orders 0:3 against the oracle's first steps.

`--impl reference` times the CPU port alone (the Julia reference cannot run here: no Julia), at the SAME N_samples
as the GPU arm of the same --gpus (2^10 per GPU), on a bounded number of steps.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ORDERS = range(0, 5)
N_TAU = 200
N_PER_GPU = 2 ** 10
METRIC = "qmc_diagram_evals_per_sec"
UNIT = "diagram_evals/s"


def workload_config(n_gpus, N):
    return {"workload": "C1 README single-orbital Anderson, Bethe bath (beta=10,U=1,eps=0.1,V=0.5), n_tau=200, "
                        "orders 0:4, orders_bare 0:4; one step = one full inchworm! run",
            "N_samples": N, "n_tau": N_TAU, "orders": "0:4", "orders_bare": "0:4", "samples_per_gpu": N_PER_GPU,
            "parallelism": "sobol-index-shard x%d, 1 all-reduce of the block sums per step" % n_gpus,
            "l2": "working set (tables+programs < 1 MB) is cache-resident by construction; 256 MiB L2 flush "
                  "write between timed runs"}


def diagram_evals(N, n_tau=N_TAU, bold_steps=None):
    """N*[sum bare (2n-1)!!] + (n_tau-2)*N*280 (+ the order-0 single evaluations)."""
    bare = 1 + N * (1 + 3 + 15 + 105)
    steps = (n_tau - 2) if bold_steps is None else bold_steps
    return bare + steps * (1 + N * 280)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed regions: NVML every ~20 ms (nvidia_ml_py), falling
    back to nvidia-smi polling.  (A 5 ms period cost the public-API figure 0.35 ms of a 3.8 ms call: the sampler is a
    Python thread and takes the interpreter lock at every wake-up; profiles/e2e_breakdown.py measures the call without it.)"""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, device_index):
        super().__init__(daemon=True)
        self.dev, self.samples, self.reasons, self.stop_flag, self.max_mhz = device_index, [], set(), False, None
        self.source = "nvidia-smi"

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[self.dev]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.dev
        h = nv.nvmlDeviceGetHandleByIndex(idx)
        self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        bits = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        self.source = "nvml"
        while not self.stop_flag:
            self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            for n, b_ in bits.items():
                if r & b_:
                    self.reasons.add(n)
            time.sleep(0.02)

    def run(self):
        try:
            return self._run_nvml()
        except Exception:
            self.source = "nvidia-smi"
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for n, v in zip(self.NAMES, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples), "source": self.source}


PARITY_TOL = 1e-10


def relerr_elem(a, b, floor=1e-3):
    """Element-wise relative difference; components below `floor` x the largest are compared with that floor."""
    a, b = np.asarray(a), np.asarray(b)
    scale = np.maximum(np.abs(b), floor * max(float(np.abs(b).max()), 1e-300))
    return float((np.abs(a - b) / scale).max())


def cpu_port_run(threads, bold_steps, N, pool=True, keep=False):
    """Bounded sample of the workload on the CPU oracle port: bare step + first `bold_steps` bold steps."""
    import models
    from oracle import oracle as orc
    ex, grid, f = models.anderson(n_tau=N_TAU)
    pl = ex.flatten()
    t0 = time.perf_counter()
    res = orc.inchworm(pl, ex.P, ORDERS, ORDERS, N, threads=threads, max_bold_steps=bold_steps, pool=pool)
    dt = time.perf_counter() - t0
    if keep:
        return res["evals"] / dt, dt, res["evals"], res
    return res["evals"] / dt, dt, res["evals"]


def run_reference(args):
    """Reference arm: the reference's own (CPU) implementation of the path.  The Julia package cannot
    run here (no Julia toolchain in the image), so the oracle port stands in, on all host cores."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    N = N_PER_GPU * max(args.gpus, 1)            # the GPU arm's N_samples at this --gpus (weak scaling)
    bold_steps = max(2, 40 // max(args.gpus, 1))  # bounded: the same CPU work per step whatever --gpus
    vals = []
    for i in range(args.warmup + args.steps):
        v, dt, ev = cpu_port_run(cores, bold_steps, N)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals]) * 1e3)
    sample = "bare step + first %d of %d bold steps at N_samples=%d, %d host threads (split_count rule)" % (
        bold_steps, N_TAU - 2, N, cores)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus, N),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "oracle port (C++ restatement of the reference algorithm, persistent worker threads = ranks); the Julia+MPI "
                "reference itself cannot run: no Julia in the image"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stress", action="store_true", help="skip the C5 stress-step section")
    ap.add_argument("--no-extra", action="store_true", help="skip the C2 / C3 / C4 sections (one GPU only)")
    ap.add_argument("--stress-log2n", type=int, default=20)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import models
    from qinchworm_b200 import lib, mpi
    from qinchworm_b200.inchworm import Solver, inchworm

    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    args.warmup = max(args.warmup, 3)
    N = N_PER_GPU * world

    # ---- set-up (untimed): model, tables, compiled entries resident on the device ----
    ex, grid, f = models.anderson(n_tau=N_TAU)
    P_atomic = ex.P.copy()
    ctx = lib.Context(device=local)
    solver = Solver(ex, ctx=ctx)
    comm_kind = mpi.init_comm(ctx) if world > 1 else "single"
    from qinchworm_b200.inchworm import MODE_BARE, _bold_entries
    bare = [solver.make_entry(MODE_BARE, o, 2 * o, N) for o in ORDERS]
    bold = _bold_entries(solver, ORDERS, N, None, None)
    bare_ids, bold_ids = [t.entry_id for t in bare], [t.entry_id for t in bold]
    stats = {t.entry_id: ctx.entry_stats(t.entry_id) for t in bare + bold}
    flops_bold_sample = sum(stats[i]["flops_per_sample"] for i in bold_ids)
    flops_deep_sample = sum(stats[t.entry_id]["flops_per_sample"] for t in bold if t.order >= 3)
    evals_per_run = diagram_evals(N)
    fp64_peak = ctx.measure_fp64_peak()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def device_run():
        ctx.set_P(0, P_atomic)
        ctx.inchworm_run(bare_ids, bold_ids, N, want_contribs=False)
        return ctx.last_device_ms()

    # ---- value: device-resident run, K steps ----
    for _ in range(args.warmup):
        device_run()
    l0 = ctx.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms = []
    for _ in range(args.steps):
        flush.fill_(1)          # L2 flush between timed runs (not inside the event bracket)
        torch.cuda.synchronize()
        dev_ms.append(device_run())
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3 / args.steps
    launches = ctx.launch_count() - l0
    ms_per_step = max_over_ranks(float(np.mean(dev_ms)))
    wall_ms = max_over_ranks(wall_ms)
    value = evals_per_run / (ms_per_step * 1e-3)
    P_dev = ctx.get_P()
    ctx.set_P(0, P_atomic)
    hist_dev = ctx.inchworm_run(bare_ids, bold_ids, N, want_contribs=True)   # untimed: per-entry contributions of every step
    assert np.array_equal(ctx.get_P(), P_dev), "device-resident run is not reproducible bit for bit"

    # ---- e2e: the public API call with host buffers (wall clock around the call) ----
    def host_run(device_resident):
        ex.P[:] = P_atomic
        t = time.perf_counter()
        inchworm(ex, grid, ORDERS, ORDERS, N, solver=solver, device_resident=device_resident)
        return (time.perf_counter() - t) * 1e3
    host_run(None)
    barrier()
    e2e_ms = float(np.mean([host_run(None) for _ in range(args.steps)]))
    barrier()
    e2e_ms = max_over_ranks(e2e_ms)
    parity = float(np.abs(ex.P - P_dev).max() / np.abs(P_dev).max())
    host_run(False)
    barrier()
    stepped_ms = max_over_ranks(float(np.mean([host_run(False) for _ in range(max(2, min(args.steps, 3)))])))
    parity_stepped = float(np.abs(ex.P - P_dev).max() / np.abs(P_dev).max())
    sampler.stop_flag = True
    sampler.join(timeout=2)
    e2e_value = evals_per_run / (e2e_ms * 1e-3)
    bs = ctx.bsize
    h2d = N_TAU * bs * 16                                        # atomic P table
    d2h = N_TAU * bs * 16 + N_TAU * (len(bare_ids) + len(bold_ids)) * bs * 16   # final P + order-resolved contributions

    # ---- roofline of the dominant kernel: one profiled run (event pair around every launch) ----
    ctx.profile_enable(True)
    ctx.profile_read(reset=True)
    device_run()
    prof = ctx.profile_read(reset=True)
    ctx.profile_enable(False)
    # the bold steps run as ONE launch of the persistent run kernel (all 198 steps, grid barrier per step) when a step is
    # too small to fill the machine, else as one step kernel per step (reduction, all-reduce and P update fused in its tail)
    n_bold_steps = N_TAU - 2
    if "run_kernel" in prof:
        dom_name, steps_per_launch = "run_kernel", n_bold_steps
        kernel_label = "scalar_run_kernel<real>: all %d bold steps of the run in one cooperative launch" % n_bold_steps
    else:
        dom_name = "step_real" if "step_real" in prof else "step_complex"
        steps_per_launch = 1
        kernel_label = "scalar_step_kernel<%s> (all bold entries, orders 0-4)" % ("real" if dom_name == "step_real" else "complex")
    dom = prof.get(dom_name, {"ms": float("nan"), "launches": 1})
    n_count = mpi.split_count(N, world)[rank]
    # average duration of the dominant kernel's launches: CUDA events around every launch of one extra (profiled) run
    dom_ms = dom["ms"] / max(dom["launches"], 1)
    if steps_per_launch == 1:
        dom_ms = ms_per_step / (N_TAU - 1)      # timed region / launches (every launch is a step kernel)
    dom_flops = flops_bold_sample * n_count * steps_per_launch      # algorithmic chain FLOPs of one launch
    achieved = dom_flops / (dom_ms * 1e-3) / 1e12
    total_prof = sum(v["ms"] for v in prof.values())
    # What the kernel EXECUTES per sample-set (counted from the lane program it runs, real arithmetic): one FP64
    # instruction and one 8-byte shared-memory operand per factor — far fewer than the complex chain of the reference
    # that `achieved` credits it with.  FP64 pipe share = executed FP64 instructions / (DFMA peak / 2 flops).
    ex_ops, ex_loads = 0.0, 0.0
    for t in bold:
        lp = ctx.entry_lane_program(t.entry_id)
        n_, K_ = t.order, lp["K"]
        for s_i, M_, n_rec, _ in lp["sections"]:
            ex_ops += n_rec * (max(n_ - 1, 0) + M_ * (K_ - 1) + (1 if n_ else 0) + M_)
            ex_loads += n_rec * (n_ + M_ * K_)
        for d_, c_ in zip(lp["segdef"], lp["seg_coef"]):
            ln = int((d_ != 0xFFFF).sum())
            ex_ops += ln - 1 + (1 if c_ != 0xFFFF else 0)
            ex_loads += ln + 1
        nD_ = lp["seg0"] - (2 * n_ + 2) * ctx.S
        ex_ops += 10 * nD_ + 8 * (2 * n_ + 2) * ctx.S           # three-point interpolation of every pair-interaction / propagator slot
    ex_inst_per_s = ex_ops * n_count * steps_per_launch / (dom_ms * 1e-3)
    ncu = {}
    for fn in ("r2_ncu_c1_run_summary.csv", "r1_ncu_c1_step_summary.csv"):
        try:
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            vals = {}
            for line in open(os.path.join(ROOT, "profiles", fn)):
                f_ = line.strip().split(",")
                if len(f_) >= 3:
                    vals[f_[0]] = (f_[1], f_[2])
            ncu = {"source": "profiles/" + fn,
                   "traffic": sum(float(vals[k][1]) * mult.get(vals[k][0], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in vals),
                   "fp64_pipe_pct": float(vals["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"][1]),
                   "smem_wavefront_pct": float(vals["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"][1]),
                   "issue_per_cycle": float(vals["smsp__issue_active.avg.per_cycle_active"][1])}
            break
        except Exception:
            continue
    traffic = ncu.get("traffic")
    traffic_src = ("%s (dram__bytes_read.sum + dram__bytes_write.sum of one launch)" % ncu["source"]) if ncu else None
    roofline = {"bound": "fp64_fma", "bound_note": "BASELINE.json asks for the fraction of FP64 peak: the path is neither HBM-bound (tables < 1 MB, dram traffic "
                "below) nor tensor-core work (blocks <= 4x4, DESIGN.md section 4).  `frac` credits the kernel with the reference's complex "
                "multiply-add chain (8 FLOPs per factor); what it executes is `executed` (real arithmetic on factorised records), and what "
                "binds it is the shared-memory operand pipe (`smem_wavefront_frac`), see DESIGN.md section 4",
                "kernel": kernel_label, "achieved": achieved,
                "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                "peak_source": "measured in this process by qiw_measure_fp64_peak (DFMA-saturating kernel); "
                               "MEASURED_PEAKS.json has no FP64 entry",
                "flops_per_launch": dom_flops, "ms_per_launch": dom_ms, "steps_per_launch": steps_per_launch,
                "us_per_step": dom_ms * 1e3 / steps_per_launch,
                "executed": {"fp64_instructions_per_launch": ex_ops * n_count * steps_per_launch,
                             "shared_operand_loads_per_launch": ex_loads * n_count * steps_per_launch,
                             "executed_frac": ex_inst_per_s / (fp64_peak * 1e12 / 2.0),
                             "smem_operand_frac": (ex_loads * n_count * steps_per_launch * 8.0 / (dom_ms * 1e-3)) / (128.0 * 148 * (sampler.summary()["sm_mhz"] or 1965.0) * 1e6),
                             "note": "counted from the lane program (configuration sums + segment products + interpolation); FP64 pipe "
                                     "peak = DFMA peak / 2 instructions; shared-memory peak = 128 B/clk/SM",
                             "ncu_fp64_pipe_frac": ncu.get("fp64_pipe_pct", float("nan")) / 100.0 if ncu else None,
                             "ncu_smem_wavefront_frac": ncu.get("smem_wavefront_pct", float("nan")) / 100.0 if ncu else None,
                             "ncu_source": ncu.get("source") if ncu else None},
                "traffic": traffic, "traffic_unit": "bytes", "traffic_source": traffic_src,
                "launches_per_run": dom["launches"],
                "kernel_share_of_step": dom["ms"] / total_prof if total_prof else None,
                "ms_per_launch_profiled": dom["ms"] / max(dom["launches"], 1),
                "profile_ms": {k: round(v["ms"], 4) for k, v in prof.items()},
                "profile_launches": {k: v["launches"] for k, v in prof.items()},
                "whole_run_frac": (flops_bold_sample * n_count * (N_TAU - 2)) / (ms_per_step * 1e-3) / 1e12 / fp64_peak}

    # ---- the same step kernel with the machine full (not the headline: the C1 step has only 2^10 samples) ----
    saturated = None
    if world == 1:
        saturated = {}
        tau = grid.tau
        for label, max_order, n_sat in (("orders_0_4_N_2^17", 4, 2 ** 17), ("orders_0_6_N_2^14", 6, 2 ** 14)):
            ent = _bold_entries(solver, range(0, max_order + 1), n_sat, None, None)
            sids = [t.entry_id for t in ent]
            sst = [ctx.entry_stats(i) for i in sids]
            fl = sum(x["flops_per_sample"] for x in sst)
            tops = sum(x["n_top"] for x in sst)
            for _ in range(2):
                ctx.eval(0.0, tau[100], tau[101], sids, n_sat)
            ms = []
            for _ in range(3):
                ctx.eval(0.0, tau[100], tau[101], sids, n_sat)
                ms.append(ctx.last_device_ms())
            m_ = float(np.median(ms))
            saturated[label] = {"ms_per_launch": m_, "diagram_evals_per_s": n_sat * tops / (m_ * 1e-3),
                                "algorithmic_tflops": fl * n_sat / (m_ * 1e-3) / 1e12,
                                "frac_of_measured_fp64_peak": fl * n_sat / (m_ * 1e-3) / 1e12 / fp64_peak}

    # ---- sector blocks larger than 1x1: one bold step of the two-band e_g model (C4), block_walk_kernel ----
    block_model = None
    if world == 1:
        ex4, grid4, _ = models.two_band(n_tau=64)
        ctx4 = lib.Context(device=local)
        solver4 = Solver(ex4, ctx=ctx4)
        n4 = 2 ** 12
        ent = _bold_entries(solver4, range(0, 4), n4, None, None)
        sids = [t.entry_id for t in ent]
        sst = [ctx4.entry_stats(i) for i in sids]
        fl = sum(x["flops_per_sample"] for x in sst)
        tops = sum(x["n_top"] for x in sst)
        ms = []
        for k in range(4):
            ctx4.eval(0.0, grid4.tau[30], grid4.tau[31], sids, n4)
            if k:
                ms.append(ctx4.last_device_ms())
        m_ = float(np.median(ms))
        block_model = {"workload": "two-band e_g model (9 sectors, blocks 1/2/4), orders 0:3, one bold step, N = 2^12",
                       "ms_per_launch": m_, "diagram_evals_per_s": n4 * tops / (m_ * 1e-3),
                       "algorithmic_tflops": fl * n4 / (m_ * 1e-3) / 1e12,
                       "frac_of_measured_fp64_peak": fl * n4 / (m_ * 1e-3) / 1e12 / fp64_peak}
        ctx4.close()

    # ---- C5 stress step (BASELINE.json configs[4]): orders 0:6, n_tau = 400, N_samples = 2^20 FIXED and sharded over
    #      the GPUs of the job (strong scaling), three bold steps in the middle of the run, through qiw_eval (one launch
    #      + all-reduce per step).  Checked against the oracle at N = 2^6 on the same entries (the oracle cannot do 2^20).
    stress = None
    if not args.no_stress:
        ex5, grid5, _ = models.anderson(n_tau=400)
        ctx5 = lib.Context(device=local)
        solver5 = Solver(ex5, ctx=ctx5)
        kind5 = mpi.init_comm(ctx5) if world > 1 else "single"
        t0 = time.perf_counter()
        N5 = 2 ** args.stress_log2n
        ent5 = _bold_entries(solver5, range(0, 7), N5, None, None)
        t_compile = time.perf_counter() - t0
        ids5 = [t.entry_id for t in ent5]
        st5 = [ctx5.entry_stats(i) for i in ids5]
        tops5 = sum(x["n_top"] for x in st5)
        fl5 = sum(x["flops_per_sample"] for x in st5)
        tau5 = grid5.tau
        # parity at N = 2^6: all 37 entries (orders 0-6) against the oracle, element-wise per entry
        got = ctx5.eval(0.0, tau5[200], tau5[201], ids5, 64)
        c5_par = None
        if rank == 0:
            from oracle import oracle as orc
            o5 = orc.Oracle(solver5.payload, ex5.P, threads=os.cpu_count() or 1)
            for j, t in enumerate(ent5):
                o5.set_topologies(j, lib.MODE_BOLD, t.order, t.n_pts_after, t.topologies[0], t.topologies[1])
            ref = o5.eval(0.0, tau5[200], tau5[201], list(range(len(ent5))), 64)
            c5_par = max(relerr_elem(got[j], ref[j]) for j in range(len(ent5)))
        barrier()                                           # rank 0 has spent seconds in the oracle: line the ranks up again
        ctx5.eval(0.0, tau5[200], tau5[201], ids5, N5)      # warm-up: fills the simplex-root cache
        barrier()
        dms = []
        t0 = time.perf_counter()
        for k in range(3):
            ctx5.eval(0.0, tau5[200 + k], tau5[201 + k], ids5, N5)
            dms.append(ctx5.last_device_ms())
        torch.cuda.synchronize()
        wall5 = max_over_ranks((time.perf_counter() - t0) / 3)
        dev5 = max_over_ranks(float(np.mean(dms)))
        stress = {"workload": "C5: Anderson orders 0:6, n_tau=400, N_samples=2^%d total (strong scaling), one bold step; mean of 3"
                              % args.stress_log2n, "n_gpus": world, "scaling": "strong", "collective": kind5,
                  "topologies": tops5, "configurations": sum(x["n_leaves"] for x in st5), "host_compile_s": t_compile,
                  "step_ms_device_max": dev5, "step_ms_wall_max": wall5,
                  "diagram_evals_per_s": N5 * tops5 / (wall5), "algorithmic_tflops": fl5 * N5 / wall5 / 1e12,
                  "algorithmic_frac_of_fp64_peak_per_gpu": fl5 * N5 / wall5 / 1e12 / fp64_peak / world,
                  "projected_full_run_s": wall5 * 398,
                  "parity_N64_vs_oracle_max_rel_elem": c5_par, "parity_pass": (c5_par is None or c5_par < PARITY_TOL)}
        ctx5.close()

    # ---- the other BASELINE.json configurations as short whole-workload runs with their own checks (one GPU) ----
    extra = None
    if world == 1 and not args.no_extra:
        from qinchworm_b200 import ppgf
        from qinchworm_b200.inchworm import correlator_2p
        extra = {}
        cores_ = os.cpu_count() or 1
        # C2 bench/bethe_gf_convergence: correlator_2p G(tau), orders_gf 0:3, all grid points in one batched launch
        ex2, grid2, _ = models.bethe_two_state(n_tau=128)
        ctx2 = lib.Context(device=local)
        solver2 = Solver(ex2, ctx=ctx2)
        inchworm(ex2, grid2, range(0, 4), range(0, 4), 2 ** 10, solver=solver2)
        tops2 = sum(len(lib.topologies(o, k, True)[1]) for o in range(0, 4) for k in ([0] if o == 0 else range(1, 2 * o)))
        c2 = {"workload": "C2 bethe_gf_convergence: spinless level on a Bethe bath, n_tau=128, inchworm orders 0:3 at N=2^10, then "
                          "correlator_2p orders 0:3 at N_samples = 2^10 / 2^13 / 2^16 (all 127 grid points per launch)", "sweep": {}}
        for n2 in (2 ** 10, 2 ** 13, 2 ** 16):
            correlator_2p(ex2, grid2, range(0, 4), n2, solver=solver2)
            t = time.perf_counter()
            g2 = correlator_2p(ex2, grid2, range(0, 4), n2, solver=solver2)[0]
            dt = time.perf_counter() - t
            c2["sweep"]["N_2^%d" % int(np.log2(n2))] = {"wall_ms": dt * 1e3, "diagram_evals_per_s": n2 * tops2 * (grid2.n_tau - 1) / dt}
            if n2 == 2 ** 10 and not args.no_cpu_baseline:
                from oracle import oracle as orc
                c2["max_rel_diff_vs_oracle_G_N_2^10"] = relerr_elem(g2, orc.correlator_2p(ex2.flatten(), ex2.P, range(0, 4), n2, threads=cores_))
        ctx2.close()
        extra["c2_correlator"] = c2
        # C3 bench/fermi_hubbard_dimer: orders 0:4, density matrix against exact diagonalisation of the dimer
        ex3, grid3, _ = models.hubbard_dimer_impurity(n_tau=64)
        ctx3 = lib.Context(device=local)
        solver3 = Solver(ex3, ctx=ctx3)
        P3 = ex3.P.copy()
        ms3 = []
        for _ in range(3):
            ex3.P[:] = P3
            t = time.perf_counter()
            inchworm(ex3, grid3, range(0, 5), range(0, 5), 2 ** 12, solver=solver3)
            ms3.append((time.perf_counter() - t) * 1e3)
        ppgf.normalize(ex3)
        rho3 = ex3.ed.to_fock_basis(ppgf.density_matrix(ex3))
        extra["c3_hubbard_dimer"] = {"workload": "C3 fermi_hubbard_dimer: orders 0:4, n_tau=64, N_samples=2^12, whole inchworm! run through the public API",
                                     "inchworm_wall_ms": float(np.median(ms3[1:])), "device_ms": ctx3.last_device_ms(),
                                     "max_abs_rho_diff_vs_exact_ed": float(np.abs(rho3 - models.hubbard_dimer_exact_rho()).max())}
        ctx3.close()
        # C4 bench/two_band_eg_model_discrete_bath: whole run at orders 0:3 (orders 0:4 compiles for 13 s: one bold step of it is
        # checked against the oracle in tests/test_gpu_parity.py::test_c4_order4_bold_step_vs_oracle)
        ex4b, grid4b, _ = models.two_band(n_tau=32)
        ctx4b = lib.Context(device=local)
        solver4b = Solver(ex4b, ctx=ctx4b)
        P4 = ex4b.P.copy()
        ms4 = []
        for _ in range(2):
            ex4b.P[:] = P4
            l0_ = ctx4b.launch_count()
            t = time.perf_counter()
            Po4, _ = inchworm(ex4b, grid4b, range(0, 4), range(0, 4), 2 ** 10, solver=solver4b)
            ms4.append((time.perf_counter() - t) * 1e3)
            n_l4 = ctx4b.launch_count() - l0_
        c4 = {"workload": "C4 two_band_eg_model: 9 sectors (blocks 1/2/4), orders 0:3, n_tau=32, N_samples=2^10, whole inchworm! run",
              "inchworm_wall_ms": ms4[-1], "device_ms": ctx4b.last_device_ms(), "launches": n_l4}
        ppgf.normalize(ex4b)
        c4["trace_rho"] = float(sum(np.trace(d).real for d in ppgf.density_matrix(ex4b)))
        if not args.no_cpu_baseline:
            from oracle import oracle as orc
            r4 = orc.inchworm(models.two_band(n_tau=32)[0].flatten(), P4, range(0, 4), range(0, 4), 2 ** 10, threads=cores_, max_bold_steps=1)
            c4["max_rel_diff_vs_oracle_first_steps"] = relerr_elem(sum(Po4.values())[1:3], sum(r4["P_orders"].values())[1:3])
        ctx4b.close()
        extra["c4_two_band"] = c4

    # ---- cold call: what a user pays the first time (fresh context, model upload, host compilation of all
    #      entries, then the run); the e2e figure above reuses the compiled session, as a production run over
    #      many inchworm! calls on one Expansion does.  Reported, never the headline; a failure here is recorded. ----
    first_call = None
    if world == 1:
        try:
            exc, gridc, _ = models.anderson(n_tau=N_TAU)
            torch.cuda.synchronize()
            t = time.perf_counter()
            ctxc = lib.Context(device=local)
            t_ctx = time.perf_counter()
            solverc = Solver(exc, ctx=ctxc)
            t_model = time.perf_counter()
            inchworm(exc, gridc, ORDERS, ORDERS, N, solver=solverc)
            t_end = time.perf_counter()
            first_call = {"total_ms": (t_end - t) * 1e3, "context_ms": (t_ctx - t) * 1e3, "model_upload_ms": (t_model - t_ctx) * 1e3,
                          "compile_and_run_ms": (t_end - t_model) * 1e3,
                          "max_rel_diff_vs_device_resident": float(np.abs(exc.P - P_dev).max() / np.abs(P_dev).max())}
            ctxc.close()
        except Exception as e:      # noqa: BLE001
            first_call = {"error": repr(e)}

    # ---- CPU baseline + parity: the oracle port runs the WHOLE C1 workload once (rank 0, N=1 GPU); its P table, Z, rho
    #      and G(tau) are the reference the GPU results are held to (element-wise, 1e-10) ----
    cpu, parity_rec = None, None
    g_dev = None
    if world == 1 and not args.no_cpu_baseline:
        # G(tau) = correlator_2p(expansion, grid, orders 0:3, N) on the converged P of the device-resident run (all ranks)
        from qinchworm_b200.expansion import add_corr_operators
        from qinchworm_b200.inchworm import correlator_2p
        ex.P[:] = P_dev
        add_corr_operators(ex, (f.c("up"), f.c_dag("up")))
        t = time.perf_counter()
        g_dev = correlator_2p(ex, grid, range(0, 4), N, solver=solver)[0]
        g_ms = (time.perf_counter() - t) * 1e3
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import oracle as orc
        cores = os.cpu_count() or 1
        if world == 1:
            v, dt, ev, res = cpu_port_run(cores, None, N, keep=True)          # the complete inchworm! run
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "seconds": dt, "inchworm_wall_ms": dt * 1e3,
                   "sample": "the whole workload once: bare step + all %d bold steps at N_samples=%d (%.3g diagram evals), "
                             "%d persistent worker threads (reference's split_count rule)" % (N_TAU - 2, N, ev, cores)}
            v_old, dt_old, _ = cpu_port_run(cores, 20, N, pool=False)
            v_new, dt_new, _ = cpu_port_run(cores, 20, N, pool=True)
            cpu["thread_per_call_vs_pool"] = {"sample": "bare + 20 bold steps", "round1_thread_per_entry_call_evals_per_s": v_old,
                                              "persistent_pool_evals_per_s": v_new}
            P_ref = res["P"]
            Zd, Zr = (1j * P_dev[-1]).sum(), (1j * P_ref[-1]).sum()
            t = time.perf_counter()
            g_ref = orc.correlator_2p(ex.flatten(), P_ref, range(0, 4), N, threads=cores)
            g_cpu_s = time.perf_counter() - t
            # G on the GPU was computed from the GPU's own P: the comparison covers inchworm! and correlator_2p end to end
            parity_rec = {"against": "CPU oracle, whole C1 workload (n_tau=200, orders 0:4, N=2^10; G: orders 0:3)",
                          "measure": "max over elements of |gpu - oracle| / max(|oracle|, 1e-3 max|oracle|)",
                          "max_rel_diff_vs_oracle_P": relerr_elem(P_dev, P_ref),
                          "max_rel_diff_vs_oracle_P_orders": max(relerr_elem(sum(hist_dev[:, j] for j, t_ in enumerate(bare + bold) if t_.order == o_),
                                                                             res["P_orders"][o_]) for o_ in ORDERS),
                          "max_rel_diff_vs_oracle_Z": float(abs(Zd - Zr) / abs(Zr)),
                          "max_rel_diff_vs_oracle_rho": relerr_elem(1j * P_dev[-1] / Zd, 1j * P_ref[-1] / Zr),
                          "max_rel_diff_vs_oracle_G": relerr_elem(g_dev, g_ref),
                          "Z": [Zd.real, Zd.imag], "rho": [float(x) for x in (1j * P_dev[-1] / Zd).real],
                          "G_gpu_ms": g_ms, "G_oracle_s": g_cpu_s, "tolerance": PARITY_TOL}
        else:
            # sharded run: the oracle at the SAME N_samples for the bare step and the first bold steps; the per-order
            # contributions of those steps are un-normalised, so they compare one to one with the device's
            n_chk = 3
            res = orc.inchworm(ex.flatten(), P_atomic, ORDERS, ORDERS, N, threads=cores, max_bold_steps=n_chk)
            worst = 0.0
            for o_ in ORDERS:
                dev_o = sum(hist_dev[:, j] for j, t_ in enumerate(bare + bold) if t_.order == o_)
                worst = max(worst, relerr_elem(dev_o[1:n_chk + 2], res["P_orders"][o_][1:n_chk + 2]))
            parity_rec = {"against": "CPU oracle at the same N_samples=%d: bare step + first %d bold steps, per-order contributions "
                                     "(all ranks hold bit-identical sums by construction; tests/multigpu_check.py checks that)" % (N, n_chk),
                          "max_rel_diff_vs_oracle_P_orders": worst, "tolerance": PARITY_TOL}
        parity_rec["pass"] = all(v < PARITY_TOL for k, v in parity_rec.items() if k.startswith("max_rel_diff"))

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(world, N),
            "inchworm_wall_ms": {"device_resident_events": ms_per_step, "device_resident_host_clock": wall_ms,
                                 "public_api_e2e": e2e_ms, "public_api_host_stepped": stepped_ms,
                                 "public_api_first_call": first_call},
            "diagram_evals_per_step": evals_per_run,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "qinchworm_b200.inchworm.inchworm(expansion, grid, orders, orders_bare, N_samples)",
                    "max_rel_diff_vs_device_resident": parity,
                    "host_stepped": {"value": evals_per_run / (stepped_ms * 1e-3), "ms": stepped_ms,
                                     "h2d_bytes_per_step": (N_TAU - 1) * N_TAU * bs * 16,
                                     "d2h_bytes_per_step": len(bare_ids) * bs * 16 + (N_TAU - 2) * len(bold_ids) * bs * 16,
                                     "max_rel_diff_vs_device_resident": parity_stepped}},
            "e2e_stepped": {"value": evals_per_run / (stepped_ms * 1e-3), "unit": UNIT, "ms": stepped_ms,
                            "api": "inchworm(..., device_resident=False): one qiw_eval per step through the three-worker seam "
                                   "(inchworm_step_bare / inchworm_step), set_ppgf!/normalize! on the host",
                            "note": "the headline e2e goes through the optional whole-run entry qiw_inchworm_run; this is the drop-in seam of INTEGRATION.md"},
            "parity": parity_rec, "stress_c5_step": stress, "other_configs": extra,
            "gpu_launches": int(launches), "collective": comm_kind, "roofline": roofline, "saturated_step_kernel": saturated, "block_model_step": block_model, "cpu_baseline": cpu, "clocks": sampler.summary()}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    extra_bad = bool(extra) and any(v > PARITY_TOL for sec in extra.values() for k, v in sec.items() if k.startswith("max_rel_diff"))
    if rank == 0 and ((parity_rec and not parity_rec["pass"]) or (stress and not stress["parity_pass"]) or extra_bad):
        raise SystemExit("bench.py: GPU results differ from the oracle by more than %g" % PARITY_TOL)


if __name__ == "__main__":
    main()
