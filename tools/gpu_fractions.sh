#!/bin/bash
# per-config utilisation of the three candidate roofs (HBM, FP32 pipe, MUFU) + issue slots, from one ncu pass per case
# (launch #3 of tools/prof_case.py).  Output: gpurun_out/fractions_<case>.csv, summarised by tools/fractions.py into profiles/rNN/fractions.json
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active
for c in ${CASES:-cfg1 cfg2 cfg2multi cfg3 cfg5 cfg2prob}; do
  timeout 300 ncu --metrics $M --clock-control none -k regex:fuse_ -s 2 -c 1 --csv --log-file gpurun_out/fractions_$c.csv python tools/prof_case.py $c 0 > /dev/null 2>&1
  echo "$c rc=$?"
done
