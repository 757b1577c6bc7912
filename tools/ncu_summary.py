#!/usr/bin/env python
"""Summarise an ncu report for profiles/:  python tools/ncu_summary.py <report.ncu-rep> <evals_in_launch> > profiles/x.txt
Needs ncu on PATH (reads the raw and source pages; works without a GPU)."""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
evals = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H, U, V = rows[0], rows[1], rows[2]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__waves_per_multiprocessor",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active"]
print(f"# ncu summary of {rep.split('/')[-1]}")
for h, u, v in zip(H, U, V):
    if h in keys or h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"):
        print(f"{h:88s} {u:16s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
ai, ii = H.index("Source"), H.index("Instructions Executed")
byop, tot = collections.Counter(), 0
for r in rows[hdr + 1:]:
    try:
        n = int(r[ii])
    except (ValueError, IndexError):
        continue
    parts = r[ai].strip().split()
    op = parts[1] if parts[0].startswith("@") else parts[0]
    byop[op.split(".")[0]] += n
    tot += n
print(f"\n# dynamic SASS mix (warp instructions executed: {tot})")
if evals:
    w = evals / 32
    print(f"# per warp-eval (32 particle-beam evaluations): {tot / w:.1f} instructions")
    for op, n in byop.most_common(24):
        print(f"{op:10s} {n / w:8.2f} per warp-eval")
