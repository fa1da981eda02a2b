#!/bin/bash
# A/B builds of the kernel TU: tools/build_variant.sh NAME [-DFLAG=..]...  ->  percept_b200/lib/variants/libpms_NAME.so
# (only the default-build kernel object is recompiled; strict kernels, api and fixtures objects are shared with libpms.so).
# Select at run time with PMS_LIB=percept_b200/lib/variants/libpms_NAME.so (read by percept_b200/__init__.py).
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/percept_b200/csrc; out=$root/percept_b200/lib/variants; obj=$root/percept_b200/lib/obj
mkdir -p $out
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v -DPMS_NS=pms_fast "$@" \
  -c $src/pms_kernels.cu -o $out/k_fast_$name.o 2> $out/k_fast_$name.ptxas.log || { cat $out/k_fast_$name.ptxas.log; exit 1; }
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libpms_$name.so $out/k_fast_$name.o $obj/k_strict.o $obj/api.o $obj/fixtures.o -ldl
echo "built $out/libpms_$name.so ($*)"
