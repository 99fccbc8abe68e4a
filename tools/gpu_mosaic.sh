#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mosaic.py -x -q -m gpu > gpurun_out/pytest_mosaic.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_mosaic.log
timeout 300 python tools/bench_mosaic.py 32 2>&1 | tee gpurun_out/bench_mosaic.txt
