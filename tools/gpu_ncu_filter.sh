#!/bin/bash
# filter tests, then ncu full capture of the filtered kernel (one launch of the bench workload)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_filter.py -x -q -m gpu > gpurun_out/pytest_filter.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_filter.log
BCMD="python bench.py --steps 3 --warmup 3 --tiles 4096 --e2e-tiles 1024 --e2e-steps 1 --no-cpu-baseline --skip-check ${BENCH_EXTRA}"
timeout 300 $BCMD > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fuse_filter -s 2 -c 1 -f -o gpurun_out/prof_filter $BCMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
