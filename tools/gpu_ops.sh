#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_scripts.py -x -q -m gpu > gpurun_out/pytest_ops.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_ops.log
timeout 600 python tools/bench_all.py 2>&1 | grep -E "confusion|upsample|mIoU" | cut -c1-330
