#!/usr/bin/env python
"""Sum warp-instructions / samples of an .ncu-rep per source-line range.  usage: ncu_regions.py rep file:lo-hi[:name] ..."""
import csv, subprocess, sys
rep = sys.argv[1]
regions = []
for a in sys.argv[2:]:
    parts = a.split(':'); f = parts[0]; lo, hi = map(int, parts[1].split('-')); name = parts[2] if len(parts) > 2 else a
    regions.append((f, lo, hi, name))
out = subprocess.run(f"ncu -i {rep} --page source --csv --print-source cuda,sass", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hd = None; agg = {}
for r in rows:
    if r and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if r and r[0] == 'Line No': hd = r; iE = hd.index('Instructions Executed'); iS = hd.index('# Samples'); continue
    if hd and len(r) > iE and r[2] == '-':
        try: n = int(r[iE]); s = int(r[iS])
        except Exception: continue
        a = agg.setdefault((cur, int(r[0])), [0, 0]); a[0] += n; a[1] += s
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print(f"total warp-instr {tot} samples {tots}")
rest = [tot, tots]
for f, lo, hi, name in regions:
    n = sum(v[0] for (ff, l), v in agg.items() if ff == f and lo <= l <= hi); s = sum(v[1] for (ff, l), v in agg.items() if ff == f and lo <= l <= hi)
    rest[0] -= n; rest[1] -= s
    print(f"{name:28s} {100*n/tot:5.1f}% instr {100*s/max(tots,1):5.1f}% samples  ({n} instr)")
print(f"{'(other)':28s} {100*rest[0]/tot:5.1f}% instr {100*rest[1]/max(tots,1):5.1f}% samples")
