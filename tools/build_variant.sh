#!/bin/bash
# tools/build_variant.sh NAME [nvcc -D flags...] — an A/B copy of libsse_b200.so with other compile-time knobs:
# cloud.jl_b200/lib/variants/libsse_b200_NAME.so (select it with SSE_B200_LIB=...; tools/ab_kernels.py does).
set -e
cd "$(dirname "$0")/../cloud.jl_b200/csrc"
name=$1; shift
NVFLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC"
mkdir -p build/var_$name ../lib/variants      # build/var_* can be deleted at any time (objects only)
for f in sse_b200 ct_kernels comm partition; do nvcc $NVFLAGS "$@" -c -o build/var_$name/$f.o $f.cu & done
wait
nvcc $NVFLAGS -shared -o ../lib/variants/libsse_b200_$name.so build/var_$name/sse_b200.o build/var_$name/ct_kernels.o build/var_$name/comm.o build/var_$name/partition.o -ldl
echo built ../lib/variants/libsse_b200_$name.so
