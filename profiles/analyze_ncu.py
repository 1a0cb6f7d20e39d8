#!/usr/bin/env python
"""Summarise an ncu report: key raw metrics per launch + executed-instruction split by CUDA source line.
usage: python profiles/analyze_ncu.py gpurun_out/prof.ncu-rep [kernel-regex] [top-n]"""
import csv, io, subprocess, sys

rep = sys.argv[1]; kre = sys.argv[2] if len(sys.argv) > 2 else "stage03"; topn = int(sys.argv[3]) if len(sys.argv) > 3 else 30
KEEP = ['Kernel Name', 'Block Size', 'Grid Size', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'launch__waves_per_multiprocessor',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum',
        'lts__t_sectors_op_read.sum', 'lts__t_sectors_op_write.sum', 'sm__inst_executed_pipe_xu.sum',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.pct', 'smsp__warps_eligible.avg.per_cycle_active']
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
print("| metric | unit | " + " | ".join(f"launch {i+1}" for i in range(len(rows) - 2)) + " |")
print("|---|---|" + "---|" * (len(rows) - 2))
for k in KEEP:
    if k in hdr:
        i = hdr.index(k)
        print(f"| {k} | {units[i]} | " + " | ".join(r[i] for r in rows[2:]) + " |")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur, agg, seen_fn = None, {}, 0
for r in rows:
    if not r: continue
    if r[0] == 'File Path':
        cur = r[1].split('/')[-1]; continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Kernel Name':
        seen_fn += 1
        if seen_fn > 1: break
        continue
    if r[0].isdigit() and len(r) > 8 and r[2] == '-':
        key = (cur, int(r[0]))
        if key in agg: continue
        agg[key] = (int(r[7]), int(r[8]), int(r[6]), r[1])
tot = sum(v[0] for v in agg.values()) or 1; tots = sum(v[2] for v in agg.values()) or 1
print(f"\nexecuted warp instructions (first launch): {tot}, stall samples: {tots}\n")
byfile = {}
for (f, l), v in agg.items():
    b = byfile.setdefault(f, [0, 0, 0]); b[0] += v[0]; b[1] += v[1]; b[2] += v[2]
print("| file | % warp inst | avg active lanes | % samples |\n|---|---|---|---|")
for f, b in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print(f"| {f} | {100*b[0]/tot:.1f} | {b[1]/max(b[0],1):.1f} | {100*b[2]/tots:.1f} |")
print(f"\ntop {topn} source lines by stall samples:\n")
print("| line | % inst | lanes | % samples | source |\n|---|---|---|---|---|")
for (f, l), v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:topn]:
    print(f"| {f}:{l} | {100*v[0]/tot:.1f} | {v[1]/max(v[0],1):.1f} | {100*v[2]/tots:.1f} | `{v[3].strip()[:100]}` |")
