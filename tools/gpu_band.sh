#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_filter.py -x -q -m gpu -k band > gpurun_out/pytest_band.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_band.log
timeout 600 python tools/bench_all.py 2>&1 | grep -E "cfg5|cfg3" | cut -c1-300
