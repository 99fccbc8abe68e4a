#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fuse.py tests/test_gpu_scripts.py -x -q -m gpu > gpurun_out/pytest_fr.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_fr.log
timeout 600 python tools/bench_all.py 2>&1 | grep -E "modeF|mIoU" | cut -c1-300
