#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + hottest CUDA source lines (needs -lineinfo and --import-source on)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum.per_cycle_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__sass_thread_inst_executed_op_dfma_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dadd_pred_on.sum', 'smsp__sass_thread_inst_executed_op_dmul_pred_on.sum',
        'smsp__inst_executed.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes_pipe_lsu_mem_local', 'smsp__average_warps_issue_stalled']
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(w) for w in want) and 'pct_of_peak_sustained_elapsed' not in h and '.per_second' not in h:
        print("%-90s %-12s %s" % (h, u, v))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur = None; agg = []; hdr = None
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 3 and r[0] == 'Line No': hdr = r; continue
    if hdr and len(r) > 7 and r[0].strip().isdigit():
        try: agg.append((int(r[6]), int(r[7]), cur, int(r[0]), r[1].strip()[:120]))
        except ValueError: pass
ts = sum(a[0] for a in agg) or 1; ti = sum(a[1] for a in agg) or 1
print("total samples", ts, "total warp-instructions", ti)
for a in sorted(agg, reverse=True)[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%d  %s" % (100.0 * a[0] / ts, 100.0 * a[1] / ti, a[2], a[3], a[4]))
