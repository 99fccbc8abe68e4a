#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_all.py > gpurun_out/bench_all.txt 2>&1; echo "bench_all rc=$?"; cut -c1-260 gpurun_out/bench_all.txt
