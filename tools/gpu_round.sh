#!/bin/bash
# One GPU-box session (round artefacts): GPU tests -> bench.py -> ncu launch list -> one full capture of the hot kernel.
# Outputs under gpurun_out/ (copied into profiles/rNN/ by hand).  Other helpers: gpu_iter.sh (tests + A/B of library builds +
# optional captures), gpu_prof.sh (one ncu capture of one case), ab.sh / variant.sh (library variants).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
BCMD="python bench.py --steps 3 --warmup 3 --tiles 4096 --e2e-tiles 1024 --e2e-steps 1 --no-cpu-baseline --skip-check --no-extra"
timeout 300 $BCMD > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $BCMD > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches rc=$?"
timeout 300 $BCMD > /dev/null 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fuse_ -s 2 -c 1 -f -o gpurun_out/prof_hot $BCMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
