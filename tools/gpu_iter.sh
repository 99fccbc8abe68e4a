#!/bin/bash
# one development iteration on the GPU: filter/fuse tests, bench_fuse A/B (IMPLS, extra libs as args), optional ncu capture (PROF_CASE, PROF_IMPL, PROF_NAME)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_filter.py tests/test_gpu_fuse.py -x -q -m gpu > gpurun_out/pytest_filter.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_filter.log
bash tools/ab.sh "${IMPLS:-0}" "$@"
if [ -n "$PROF_CASE" ]; then
  for c in $PROF_CASE; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:fuse_ -s 2 -c 1 -f -o gpurun_out/${PROF_NAME:-prof}_$c python tools/prof_case.py $c ${PROF_IMPL:-0} > gpurun_out/${PROF_NAME:-prof}_$c.log 2>&1; echo "ncu $c rc=$?"
  done
fi
