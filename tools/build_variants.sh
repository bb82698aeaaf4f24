#!/usr/bin/env bash
# Dev: compile-time variants of one kernel file (default mnv_render.cu; SRC=mnv_mlp.cu for the MLP) into
# build/variants/libmnv_b200_<name>.so (git-ignored; they travel with gpurun).
# Usage: [SRC=mnv_mlp.cu] tools/build_variants.sh name1="-DMNV_LAZY_EMPTY=1" name2="-DMNV_TRACK_REGS=1 -DMNV_UNROLL_ANCHOR=2" ...
set -euo pipefail
cd "$(dirname "$0")/../mega-nerf-viewer_b200/csrc"
OUT=../../build/variants
mkdir -p "$OUT"
ARCH="-gencode arch=compute_100a,code=sm_100a"
SRC="${SRC:-mnv_render.cu}"; BASE="${SRC%.cu}"
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  nvcc -std=c++17 -O3 $ARCH -lineinfo -Xcompiler -fPIC $flags -c "$SRC" -o "$OUT/${BASE}_$name.o"
  objs=$(ls *.o viewer/*.o | grep -v "^$BASE.o\$" | tr '\n' ' ')
  nvcc $ARCH -shared -o "$OUT/libmnv_b200_$name.so" "$OUT/${BASE}_$name.o" $objs -lcudart -lz
  echo "built $OUT/libmnv_b200_$name.so ($flags)"
done
