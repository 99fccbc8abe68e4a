#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_scripts.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python tools/bench_all.py 2>&1 | grep -E "big-mask|Error|error|Traceback" -A3 | cut -c1-330
