#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + per-source-line hot spots (needs -lineinfo builds and --import-source on)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.sum", "smsp__inst_executed_op_shared_atom.sum"]
for h, u, v in zip(hdr, units, vals):
    if h in keep or ("smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and float(v or 0) > 0.2):
        print(f"{h},{u},{v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
cur, agg = None, []
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] in ("Function Name", "Line No"): continue
    if r[0].isdigit() and len(r) > 7 and r[2] == "-":
        s = int(r[4]) if r[4].isdigit() else 0
        i = int(r[7]) if r[7].isdigit() else 0
        agg.append((cur, int(r[0]), r[1].strip()[:100], s, i))
ts, ti = sum(a[3] for a in agg) or 1, sum(a[4] for a in agg) or 1
print("# top source lines by stall samples")
for a in sorted(agg, key=lambda x: -x[3])[:top]:
    print(f"{a[0]}:{a[1]},samples {100*a[3]/ts:.1f}%,inst {100*a[4]/ti:.1f}%,{a[2]}")
print("# top source lines by executed instructions")
for a in sorted(agg, key=lambda x: -x[4])[:top]:
    print(f"{a[0]}:{a[1]},samples {100*a[3]/ts:.1f}%,inst {100*a[4]/ti:.1f}%,{a[2]}")
