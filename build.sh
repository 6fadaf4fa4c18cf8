#!/bin/bash
# Builds libplonk_b200.so (sm_100a) in-tree; translation units compile in parallel.
set -e
cd "$(dirname "$0")"
SRC=${PB200_SRC:-plonk_b200/csrc}
OUT=${PB200_OUT:-plonk_b200/libplonk_b200.so}
OBJ=${PB200_OBJ:-build/obj}
mkdir -p $OBJ
NVFLAGS="-std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -O2 $*"
pids=()
for f in capi ntt msm prover ecntt; do
  nvcc $NVFLAGS -c -o $OBJ/$f.o $SRC/$f.cu &
  pids+=($!)
done
g++ -std=c++17 -O2 -fPIC -c -o $OBJ/host_field.o $SRC/host_field.cpp
g++ -std=c++17 -O2 -fPIC -c -o $OBJ/composer.o $SRC/composer.cpp
for p in "${pids[@]}"; do wait $p; done
nvcc -shared -o "$OUT" $OBJ/capi.o $OBJ/ntt.o $OBJ/msm.o $OBJ/prover.o $OBJ/ecntt.o $OBJ/host_field.o $OBJ/composer.o
echo "built $OUT"
