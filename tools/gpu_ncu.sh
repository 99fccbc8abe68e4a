#!/bin/bash
# ncu full capture of the top kernel (one launch) + launch list
mkdir -p gpurun_out
BCMD="python bench.py --steps 3 --warmup 3 --tiles 4096 --e2e-tiles 1024 --e2e-steps 1 --no-cpu-baseline --skip-check"
timeout 300 $BCMD > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fuse_stream -s 2 -c 1 -f -o gpurun_out/prof_fuse $BCMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
