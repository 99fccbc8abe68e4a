#!/bin/bash
# usage: gpu_prof.sh case impl outname   -- ncu full capture of the 3rd launch of one fusion case
mkdir -p gpurun_out
timeout 300 python tools/prof_case.py $1 $2 > /dev/null 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fuse_ -s 2 -c 1 -f -o gpurun_out/$3 python tools/prof_case.py $1 $2 > gpurun_out/$3.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/$3.log
