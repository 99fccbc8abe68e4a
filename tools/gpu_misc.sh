#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fuse.py -x -q -m gpu > gpurun_out/pytest_misc.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_misc.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python tools/bench_all.py 2>&1 | grep -E "PROB" | cut -c1-300
