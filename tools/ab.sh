#!/bin/bash
# A/B of library builds: tools/ab.sh "impls" lib1 lib2 ...   (default build first)
mkdir -p gpurun_out; : > gpurun_out/ab.txt
IMPLS=$1; shift
for lib in default "$@"; do
  echo "== $lib" | tee -a gpurun_out/ab.txt
  if [ "$lib" = default ]; then unset PISTOSEG_B200_LIB; else export PISTOSEG_B200_LIB=$PWD/$lib; fi
  timeout 300 python tools/bench_fuse.py $IMPLS 2>&1 | tee -a gpurun_out/ab.txt
done
