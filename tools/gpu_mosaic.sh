#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mosaic.py -x -q -m gpu > gpurun_out/pytest_mosaic.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_mosaic.log
timeout 600 python tools/bench_all.py > gpurun_out/bench_all.txt 2>&1; echo "bench_all rc=$?"; tail -12 gpurun_out/bench_all.txt | cut -c1-400
