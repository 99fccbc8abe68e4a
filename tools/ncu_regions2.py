#!/usr/bin/env python
"""Per-region breakdown of an .ncu-rep of fuse_static_kernel: warp instructions and stall samples per source region.
usage: ncu_regions2.py rep"""
import csv, subprocess, sys, re
rep = sys.argv[1]
src = open('pistoseg_b200/csrc/fuse_static.cuh').read().split('\n')
def find(pat, start=0):
    for i in range(start, len(src)):
        if pat in src[i]: return i + 1
    raise KeyError(pat)
marks = [("head/tables", 1), ("load_h", find("void static_load_h")), ("rows", find("unsigned int static_rows(")), ("recheck", find("unsigned int static_recheck4")), ("prepass", find("float static_prepass")),
         ("export", find("float static_div_slow")), ("kernel setup", find("void __launch_bounds__")), ("producer", find("===== producer warp")),
         ("export dispatch", find("auto export_units")), ("tile head", find("const bool is_export")), ("exact pass", find("auto exact_warp")),
         ("vector pass", find("// ---- vector pass")), ("tile tail", find("}  // !is_export")), ("duo", find("constexpr int kDThreads")), ("host", find("// host side"))]
out = subprocess.run(f"ncu -i {rep} --page source --csv --print-source cuda,sass", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hd = None; agg = {}
for r in rows:
    if r and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hd = r; iE = hd.index('Instructions Executed'); iS = hd.index('# Samples'); continue
    if hd and len(r) > iE and r[2] == '-':
        try: n = int(r[iE]); s = int(r[iS]); ln = int(r[0])
        except Exception: continue
        if cur == 'fuse_static.cuh':
            reg = [m[0] for m in marks if m[1] <= ln][-1]
        else:
            reg = cur
        a = agg.setdefault(reg, [0, 0]); a[0] += n; a[1] += s
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print(f"total warp-instr {ti}, samples {ts}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{100*v[0]/ti:6.1f}% instr {100*v[1]/max(ts,1):6.1f}% smp   {k}")
