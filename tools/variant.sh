#!/bin/bash
# tools/variant.sh NAME "EXTRA nvcc flags" file.cu [file.cu ...]: variants/lib_NAME.so = the default build with the listed sources recompiled with EXTRA
set -e
name=$1; extra=$2; shift 2
mkdir -p variants build_$name
objs=$(ls build/*.o)
for f in "$@"; do
  b=$(basename $f .cu)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude -Ipistoseg_b200/csrc -diag-suppress 128 $extra -c $f -o build_$name/$b.o &
  objs=$(echo "$objs" | grep -v "build/$b.o")
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/lib_$name.so $objs build_$name/*.o
echo "variants/lib_$name.so"
