#!/usr/bin/env python
"""gpurun_out/fractions_<case>.csv (tools/gpu_fractions.sh) -> profiles/<round>/fractions.json: per BASELINE config the kernel that ran, its
duration under ncu and the utilisation of each candidate roof: hbm (DRAM throughput), fp32 (FMA pipe), mufu (XU pipe),
plus issue slots / ALU pipe / shared-memory wavefronts, with the largest one named.  usage: fractions.py profiles/r02"""
import csv, glob, json, os, sys
out_dir = sys.argv[1]
res = {}
for path in sorted(glob.glob("gpurun_out/fractions_*.csv")):
    case = os.path.basename(path)[len("fractions_"):-4]
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    if not rows: continue
    m = {r[12]: float(r[14].replace(",", "")) for r in rows}
    u = {r[12]: r[13] for r in rows}
    dram = sum(m[k] * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    f = {"hbm": m["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"] / 100,
         "fp32": m["sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"] / 100,
         "mufu": m["sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active"] / 100,
         "issue": m["smsp__issue_active.avg.pct_of_peak_sustained_active"] / 100,
         "alu": m["sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"] / 100,
         "smem_wavefronts": m["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"] / 100}
    res[case] = {"kernel": rows[0][4], "duration_us_under_ncu": m["gpu__time_duration.sum"] * {"us": 1, "ms": 1e3, "ns": 1e-3}.get(u["gpu__time_duration.sum"], 1),
                 "dram_bytes": dram, "warps_per_sm": m["sm__warps_active.avg.pct_of_peak_sustained_active"] / 100 * 64,
                 "fractions": {k: round(v, 4) for k, v in f.items()}, "largest": max(f, key=f.get)}
json.dump(res, open(os.path.join(out_dir, "fractions.json"), "w"), indent=1)
print(json.dumps(res, indent=1))
