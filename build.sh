#!/bin/bash
# Builds libplonk_b200.so (sm_100a) in-tree.  Usage: ./build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
SRC=plonk_b200/csrc
OUT=plonk_b200/libplonk_b200.so
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a \
     -Xcompiler -fPIC -Xcompiler -O2 -shared "$@" \
     -o "$OUT" $SRC/capi.cu $SRC/ntt.cu $SRC/msm.cu $SRC/host_field.cpp $(ls $SRC/prover*.cu 2>/dev/null) -lcudart
echo "built $OUT"
