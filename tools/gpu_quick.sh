#!/bin/bash
# quick A/B: filter tests + bench_fuse of the default build and any extra libs given as arguments
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_filter.py tests/test_gpu_fuse.py -x -q -m gpu > gpurun_out/pytest_filter.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_filter.log
bash tools/ab.sh "${IMPLS:-0}" "$@"
