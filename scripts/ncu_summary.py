#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) and optionally its source page into text for profiles/.
usage: python scripts/ncu_summary.py gpurun_out/prof.ncu-rep [rows_per_launch]"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
rows_per_launch = int(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, body = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
for k, b in enumerate(body):
    print(f"--- launch {k} of {rep}")
    for wname in want:
        if wname in hdr:
            i = hdr.index(wname)
            print(f"{wname:75s} {b[i]} {units[i]}")
    if rows_per_launch and "smsp__inst_executed.sum" in hdr:
        inst = float(b[hdr.index("smsp__inst_executed.sum")])
        print(f"{'thread-instructions per decision (32 x warp inst / rows)':75s} {32 * inst / rows_per_launch:.1f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(srows) if r and r[0] == "Address"]
if starts:
    h = srows[starts[0]]
    end = starts[1] - 1 if len(starts) > 1 else len(srows)
    b = [r for r in srows[starts[0] + 1:end] if len(r) == len(h)]
    ia, ie, iss = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
    mix, stall = Counter(), Counter()
    tot = 0
    for r in b:
        t = r[ia].split()
        op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
        mix[op] += int(r[ie])
        stall[op] += int(r[iss])
        tot += int(r[ie])
    print(f"--- SASS mix, first launch (warp instructions, share, stall samples); total {tot}")
    for op, n in mix.most_common(16):
        print(f"{op:10s} {n:12d} {100 * n / tot:5.1f}%  stalls {stall[op]}")
    print("--- top stall sites")
    for r in sorted(b, key=lambda r: -int(r[iss]))[:10]:
        print(f"{r[iss]:>6s}  {r[ia][:100]}")
