#!/usr/bin/env python
"""Condensed summary of one kernel of an ncu report (the numbers DESIGN.md / bench.py quote).
usage: ncu_summary.py report.ncu-rep [row_index]"""
import csv, subprocess, sys
rep = sys.argv[1]
row = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2 + row]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
        "smsp__sass_inst_executed_op_shared_st.sum", "smsp__sass_inst_executed_op_global_ld.sum", "l1tex__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "sm__sass_thread_inst_executed_op_dmul_pred_on.sum", "sm__sass_thread_inst_executed_op_dadd_pred_on.sum"]
print("metric,unit,value")
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print("%s,%s,%s" % (w, units[i], r[i]))
for i, h in enumerate(hdr):
    if ("dmma" in h or "pipe_tensor" in h) and h not in want and r[i] not in ("", "0", "n/a"):
        print("%s,%s,%s" % (h, units[i], r[i]))
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
        if float(r[i] or 0) >= 0.2:
            print("%s,%s,%s" % (h, units[i], r[i]))
