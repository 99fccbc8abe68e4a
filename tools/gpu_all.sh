#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python tools/bench_fuse.py 0 > gpurun_out/bench_fuse.txt 2>&1; cat gpurun_out/bench_fuse.txt
timeout 600 python tools/bench_all.py > gpurun_out/bench_all.txt 2>&1; echo "bench_all rc=$?"; cut -c1-200 gpurun_out/bench_all.txt | grep -v "^$"
