#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_filter.py -x -q -m gpu > gpurun_out/pytest_filter.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_filter.log
timeout 300 python tools/bench_fuse.py ${IMPLS:-2,4} > gpurun_out/bench_fuse.txt 2>&1; echo "bench rc=$?"; cat gpurun_out/bench_fuse.txt
