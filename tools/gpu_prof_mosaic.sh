#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_mosaic.py 4 > /dev/null 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mosaic_kernel -s 2 -c 1 -f -o gpurun_out/prof_mosaic2 python tools/bench_mosaic.py 4 > gpurun_out/prof_mosaic2.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/prof_mosaic2.log
