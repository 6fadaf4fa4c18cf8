#!/bin/bash
# Builds plonk_b200/libplonk_b200_patched.so from a scratch copy of csrc/ with every tools/patches/*.patch
# applied (CPU only; run before tools/ab_patches.sh goes to the GPU).  The default library is untouched.
set -e
cd "$(dirname "$0")/.."
rm -rf build/patched_src && mkdir -p build/patched_src/plonk_b200 build/patched_src/include
cp -r plonk_b200/csrc build/patched_src/plonk_b200/csrc
cp include/*.h include/*.hpp build/patched_src/include/ 2>/dev/null || true
for p in tools/patches/quot4n_prover.patch tools/patches/lagrange_wires_prover.patch; do
  (cd build/patched_src && patch -p1 --quiet -i "../../$p")
  echo "applied $p"
done
PB200_SRC=build/patched_src/plonk_b200/csrc PB200_OBJ=build/obj_patched PB200_OUT=plonk_b200/libplonk_b200_patched.so ./build.sh "$@"
