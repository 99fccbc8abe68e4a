#!/usr/bin/env python
"""Where the stall samples of an .ncu-rep sit: per stall reason, the source lines (file:line) with the most samples.
usage: ncu_stalls.py rep [reasons, comma separated] [n_lines]"""
import csv, subprocess, sys
from collections import defaultdict
rep = sys.argv[1]
reasons = (sys.argv[2] if len(sys.argv) > 2 else "long_sb,barrier,wait,no_inst,short_sb,math").split(",")
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 8
out = subprocess.run(f"ncu -i {rep} --page source --csv --print-source cuda,sass", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hd = None; agg = defaultdict(lambda: defaultdict(int)); text = {}
for r in rows:
    if r and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No' and len(r) > 3: hd = r; idx = {x: hd.index('stall_' + x) for x in reasons}; continue
    if hd and len(r) > max(idx.values()) and r[2] == '-':
        key = (cur, r[0]); text[key] = r[1].strip()[:90]
        for x, i in idx.items():
            try: agg[x][key] += int(r[i])
            except ValueError: pass
for x in reasons:
    tot = sum(agg[x].values())
    print(f"== stall_{x}: {tot} samples")
    for key, v in sorted(agg[x].items(), key=lambda kv: -kv[1])[:topn]:
        print(f"  {100 * v / max(tot, 1):5.1f}%  {key[0]}:{key[1]}  {text[key]}")
