#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics, stall reasons, opcode mix, hottest source lines.  usage: ncu_report.py rep [n_lines]"""
import csv, subprocess, sys
from collections import Counter
rep = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(f"ncu -i {rep} --page raw --csv", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr, units, vals = rows[0], rows[1], rows[2]
keys = ['Kernel Name','gpu__time_duration.sum', 'dram__bytes_read.sum ', 'dram__bytes_write.sum ', 'launch__registers_per_thread ', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum ', 'sm__inst_executed_pipe_fma.sum.pct', 'sm__inst_executed_pipe_fmaheavy.sum.pct','sm__inst_executed_pipe_alu.sum.pct', 'sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct', 'smsp__issue_active.avg.pct', 'sm__warps_active.avg.pct', 'launch__shared_mem_per_block_dynamic','sm__pipe_fma_cycles_active.avg.pct','sm__pipe_fmaheavy_cycles_active','sm__pipe_alu_cycles_active.avg.pct',
        'sm__cycles_elapsed.avg ', 'smsp__warps_eligible.avg.per_cycle_active','sass__inst_executed_local','sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active']
for i, h in enumerate(hdr):
    hh = h + ' '
    if any(k in hh for k in keys): print(f"{h} [{units[i]}] = {vals[i]}")
print("-- stalls (warps per issue-active cycle)")
st = [(float(vals[i]), h) for i, h in enumerate(hdr) if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and 'not_issued' not in h]
for v, h in sorted(st, reverse=True)[:10]: print(f"  {v:6.3f} {h.split('issue_stalled_')[1].split('_per_issue')[0]}")
out = subprocess.run(f"ncu -i {rep} --page source --csv --print-source cuda,sass", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; agg = {}; hd = None; ops = Counter(); tot_sass = 0
for r in rows:
    if r and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Function Name': continue
    if r and r[0] == 'Line No': hd = r; iE = hd.index('Instructions Executed'); iS = hd.index('# Samples'); continue
    if hd and len(r) > iE:
        try: n = int(r[iE]); s = int(r[iS])
        except Exception: continue
        if r[2] == '-':
            key = (cur, int(r[0]), r[1].strip()[:100]); a = agg.get(key, [0, 0]); a[0] += n; a[1] += s; agg[key] = a
        else:
            t = r[3].split(); op = (t[1] if t[0].startswith('@') else t[0]).split('.')[0]; ops[op] += n; tot_sass += n
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print(f"-- opcode mix (warp instr executed, total {tot_sass})")
print("  " + "  ".join(f"{op}:{100*n/max(tot_sass,1):.1f}%" for op, n in ops.most_common(22)))
print(f"-- hottest source lines (total warp-instr {tot}, samples {tots})")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:topn]:
    print(f"{100*v[0]/tot:5.1f}% instr {100*v[1]/max(tots,1):5.1f}% smp  {k[0]}:{k[1]}  {k[2]}")
print("-- most-sampled lines")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"{100*v[1]/max(tots,1):5.1f}% smp {100*v[0]/tot:5.1f}% instr  {k[0]}:{k[1]}  {k[2]}")
