#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` capture of the hot kernel: DRAM bytes per launch (bench.py's roofline.traffic)
and the utilisation of the resources that actually bind the kernel.  usage: ncu_traffic.py rep tiles_per_launch source_note"""
import csv, json, os, subprocess, sys
rep, tiles, note = sys.argv[1], int(sys.argv[2]), sys.argv[3]
out = subprocess.run(f"ncu -i {rep} --page raw --csv", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines())); hdr, units, vals = rows[0], rows[1], rows[2]
def get(name):
    i = hdr.index(name); v = float(vals[i]); u = units[i]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
res = {
    "issue_slots_frac": get("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100,
    "shared_memory_wavefronts_frac": get("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") / 100,
    "fma_pipe_frac": get("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active") / 100,
    "alu_pipe_frac": get("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active") / 100,
    "mufu_pipe_frac": get("sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active") / 100,
    "dram_throughput_frac_under_ncu": get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed") / 100,
    "warps_per_sm": get("sm__warps_active.avg.pct_of_peak_sustained_active") / 100 * 64,
    "binding": "instruction issue / dependent-instruction latency at 15 warps per SM (13 compute + producer + fixer; 128 registers x 480 threads, 202 KB shared memory: one CTA); see DESIGN.md 4.1",
    "source": note,
}
tj = {"kernel": vals[hdr.index("Kernel Name")], "tiles_per_launch": tiles, "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
      "source": note, "algorithmic_bytes_per_tile": 171440, "measured_bytes_per_tile": (rd + wr) / tiles, "other_resources": res}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json")
json.dump(tj, open(path, "w"), indent=1)
print(json.dumps(tj, indent=1))
